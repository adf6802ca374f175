"""ctypes binding of oracle/liboracle.so (the C oracle).  Test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# numpy mirror of `ibft_sig_item` (include/ibft_verify.h)
ITEM_DTYPE = np.dtype([
    ("r", "u1", 32), ("s", "u1", 32), ("digest", "u1", 32), ("signer", "u1", 20),
    ("v", "u1"), ("kind", "u1"), ("group", "<u2"), ("payload_off", "<u4"), ("payload_len", "<u4"),
])
assert ITEM_DTYPE.itemsize == 128


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "c", "ibft_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_verify_batch.restype = ctypes.c_int
    return _LIB


def _buf(b: bytes):
    return (ctypes.c_uint8 * len(b)).from_buffer_copy(b)


def keccak256(data: bytes) -> bytes:
    out = (ctypes.c_uint8 * 32)()
    lib().oracle_keccak256(_buf(data) if data else None, ctypes.c_size_t(len(data)), out)
    return bytes(out)


def ecrecover_address(digest: bytes, sig: bytes):
    if len(digest) != 32 or len(sig) != 65:
        return None
    out = (ctypes.c_uint8 * 20)()
    ok = lib().oracle_ecrecover_address(_buf(digest), _buf(sig[:32]), _buf(sig[32:64]), ctypes.c_uint8(sig[64]), out)
    return bytes(out) if ok else None


def ecrecover_pubkey(digest: bytes, sig: bytes):
    out = (ctypes.c_uint8 * 64)()
    ok = lib().oracle_ecrecover_pubkey(_buf(digest), _buf(sig[:32]), _buf(sig[32:64]), ctypes.c_uint8(sig[64]), out)
    return bytes(out) if ok else None


def pubkey_from_scalar(k: int):
    out = (ctypes.c_uint8 * 64)()
    ok = lib().oracle_pubkey_from_scalar(_buf(k.to_bytes(32, "big")), out)
    return bytes(out) if ok else None


def ecmult2(a: int, b: int, p_xy: bytes | None):
    out = (ctypes.c_uint8 * 64)()
    ok = lib().oracle_ecmult2(_buf(a.to_bytes(32, "big")), _buf(b.to_bytes(32, "big")), _buf(p_xy) if p_xy else None, out)
    return bytes(out) if ok else None


def sign_with_k(d: int, digest: bytes, k: int, low_s: bool = True):
    out = (ctypes.c_uint8 * 65)()
    ok = lib().oracle_sign_with_k(_buf(d.to_bytes(32, "big")), _buf(digest), _buf(k.to_bytes(32, "big")), int(low_s), out)
    return bytes(out) if ok else None


def _op2(name, a: int, b: int) -> int:
    out = (ctypes.c_uint8 * 32)()
    getattr(lib(), name)(_buf(a.to_bytes(32, "big")), _buf(b.to_bytes(32, "big")), out)
    return int.from_bytes(bytes(out), "big")


def _op1(name, a: int) -> int:
    out = (ctypes.c_uint8 * 32)()
    getattr(lib(), name)(_buf(a.to_bytes(32, "big")), out)
    return int.from_bytes(bytes(out), "big")


def fp_mul(a, b): return _op2("oracle_fp_mul", a, b)
def fn_mul(a, b): return _op2("oracle_fn_mul", a, b)
def fp_inv(a): return _op1("oracle_fp_inv", a)
def fn_inv(a): return _op1("oracle_fn_inv", a)


def verify_batch(items: np.ndarray, arena: bytes = b"", tables=None, group_table=None, n_threads: int = 1) -> np.ndarray:
    """items: ITEM_DTYPE array.  tables: list of (n x 20) uint8 arrays.  group_table: list of table index per group
    (0xFFFF = none).  Returns the verdict bitmap as uint32 words."""
    items = np.ascontiguousarray(items)
    n = len(items)
    bitmap = np.zeros((n + 31) // 32, dtype=np.uint32)
    arena_np = np.frombuffer(arena, dtype=np.uint8) if arena else np.zeros(1, dtype=np.uint8)
    tables = tables or []
    tabs = [np.ascontiguousarray(t, dtype=np.uint8).reshape(-1, 20) for t in tables]
    tab_ptrs = (ctypes.c_void_p * max(1, len(tabs)))(*[t.ctypes.data for t in tabs])
    tab_n = (ctypes.c_uint32 * max(1, len(tabs)))(*[len(t) for t in tabs])
    if group_table is not None:
        gt = np.ascontiguousarray(group_table, dtype=np.uint16)
        gt_ptr, n_groups = gt.ctypes.data_as(ctypes.c_void_p), len(gt)
    else:
        gt_ptr, n_groups = None, 0
    rc = lib().oracle_verify_batch(items.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(n),
                                   arena_np.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(arena)),
                                   tab_ptrs, tab_n, gt_ptr, ctypes.c_uint32(n_groups), ctypes.c_int(n_threads),
                                   bitmap.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError("oracle_verify_batch failed")
    return bitmap


# ---- bulk workload generation (tests/workloads.py: full-size configs 4 and 5)
def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


def privkeys(seed: int, n: int, first: int = 0) -> np.ndarray:
    """(n, 32) big-endian private keys of SURVEY.md §8d's derivation."""
    out = np.zeros((n, 32), dtype=np.uint8)
    lib().oracle_privkeys(ctypes.c_uint32(seed), ctypes.c_uint32(first), ctypes.c_uint32(n), out.ctypes.data_as(ctypes.c_void_p))
    return out


def addresses(privs: np.ndarray, n_threads: int = 8) -> np.ndarray:
    privs = _u8(privs).reshape(-1, 32)
    out = np.zeros((len(privs), 20), dtype=np.uint8)
    lib().oracle_addresses(privs.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(privs)), out.ctypes.data_as(ctypes.c_void_p),
                           ctypes.c_int(n_threads))
    return out


def sign_derived_batch(privs: np.ndarray, digests: np.ndarray, n_threads: int = 8) -> np.ndarray:
    """(n, 65) signatures R||S||V with the deterministic Keccak-derived nonce (k = Keccak-256(d || z || ctr)), low-s."""
    privs, digests = _u8(privs).reshape(-1, 32), _u8(digests).reshape(-1, 32)
    assert len(privs) == len(digests)
    out = np.zeros((len(privs), 65), dtype=np.uint8)
    lib().oracle_sign_derived_batch(privs.ctypes.data_as(ctypes.c_void_p), digests.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.c_uint32(len(privs)), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n_threads))
    return out


def keccak256_batch(arena: bytes | np.ndarray, offs, lens, n_threads: int = 8) -> np.ndarray:
    a = np.frombuffer(arena, dtype=np.uint8) if isinstance(arena, (bytes, bytearray)) else _u8(arena)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    out = np.zeros((len(offs), 32), dtype=np.uint8)
    lib().oracle_keccak256_batch(a.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p), lens.ctypes.data_as(ctypes.c_void_p),
                                 ctypes.c_uint32(len(offs)), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n_threads))
    return out


# ---- second CPU arm: OpenSSL 3 point arithmetic (oracle/c/ossl_recover.c); None when libcrypto was not available at build time
_OSSL = None


def ossl_lib():
    global _OSSL
    if _OSSL is None:
        build()
        so = os.path.join(_HERE, "liboracle_ossl.so")
        if not os.path.exists(so):
            subprocess.call(["make", "-C", _HERE, "-s", "liboracle_ossl.so"])
        _OSSL = ctypes.CDLL(so) if os.path.exists(so) else False
    return _OSSL or None


def ossl_verify_batch(items: np.ndarray, arena: bytes = b"", table=None, n_threads: int = 1):
    L = ossl_lib()
    if L is None:
        return None
    items = np.ascontiguousarray(items)
    n = len(items)
    bitmap = np.zeros((n + 31) // 32, dtype=np.uint32)
    arena_np = np.frombuffer(arena, dtype=np.uint8) if arena else np.zeros(1, dtype=np.uint8)
    tab = None if table is None else np.ascontiguousarray(table, dtype=np.uint8).reshape(-1, 20)
    L.ossl_verify_batch(items.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(n), arena_np.ctypes.data_as(ctypes.c_void_p),
                        ctypes.c_size_t(len(arena)), None if tab is None else tab.ctypes.data_as(ctypes.c_void_p),
                        ctypes.c_uint32(0 if tab is None else len(tab)), ctypes.c_int(n_threads), bitmap.ctypes.data_as(ctypes.c_void_p))
    return bitmap


# ---- tuned CPU arm: GLV + wNAF + binary inversions (oracle/c/fast_recover.c); bench.py's fastest CPU baseline, cross-checked
# against the plain port in tests/test_oracle_crypto.py
_FAST = None


def fast_lib():
    global _FAST
    if _FAST is None:
        build()
        so = os.path.join(_HERE, "liboracle_fast.so")
        src = os.path.join(_HERE, "c", "fast_recover.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle_fast.so"])
        _FAST = ctypes.CDLL(so)
        _FAST.fast_ecrecover_address.restype = ctypes.c_int
    return _FAST


def fast_ecrecover_address(digest: bytes, r: bytes, s: bytes, v: int):
    """address (20 bytes) or None -- same contract as ecrecover_address of the plain port"""
    out = ctypes.create_string_buffer(20)
    ok = fast_lib().fast_ecrecover_address(digest, r, s, ctypes.c_uint8(v), out)
    return out.raw if ok else None


def fast_verify_batch(items: np.ndarray, arena: bytes = b"", table=None, n_threads: int = 1) -> np.ndarray:
    L = fast_lib()
    items = np.ascontiguousarray(items)
    n = len(items)
    bitmap = np.zeros((n + 31) // 32, dtype=np.uint32)
    arena_np = np.frombuffer(arena, dtype=np.uint8) if arena else np.zeros(1, dtype=np.uint8)
    tab = None if table is None else np.ascontiguousarray(table, dtype=np.uint8).reshape(-1, 20)
    L.fast_verify_batch(items.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(n), arena_np.ctypes.data_as(ctypes.c_void_p),
                        ctypes.c_size_t(len(arena)), None if tab is None else tab.ctypes.data_as(ctypes.c_void_p),
                        ctypes.c_uint32(0 if tab is None else len(tab)), ctypes.c_int(n_threads), bitmap.ctypes.data_as(ctypes.c_void_p))
    return bitmap
