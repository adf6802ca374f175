"""The C-ABI library loads and exports every symbol include/ibft_verify.h declares (no compute without a GPU)."""
import os
import re

import pytest

import ibft_b200 as ib
from conftest import ROOT, has_gpu


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ibft_verify.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ibft_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = ib.load_library()
    syms = declared_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), s
    from importlib import import_module
    assert sorted(import_module("go-ibft_b200.engine").EXPORTS) == syms


def test_abi_version_and_struct_sizes():
    lib = ib.load_library()
    assert lib.ibft_abi_version() == 2
    assert ib.ITEM_DTYPE.itemsize == 128 and ib.GROUP_DTYPE.itemsize == 16 and ib.RESULT_DTYPE.itemsize == 56


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a CPU-only box")
def test_engine_fails_loudly_without_gpu():
    with pytest.raises(ib.EngineError) as ei:
        ib.Engine()
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
