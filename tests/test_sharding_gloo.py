"""N>1 path on CPU: world_size-2 gloo run of the sharding logic (shard bounds, padded bitmap all-gather, quorum from the
complete bitmap).  Verdicts of each shard come from the oracle here (no GPU); the GPU run of the same code path is bench.py
--gpus N and tests/test_gpu_verify.py::test_device_resident_and_sharded_path."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sharding = importlib.import_module("go-ibft_b200.sharding")
HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_bounds_cover_and_align():
    for n in (0, 1, 31, 32, 33, 3000, 10_000, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            for lo, hi in spans:
                assert (lo % 32 == 0 or lo == hi == n) and (hi % 32 == 0 or hi == n)   # empty trailing shards sit at n
                assert (hi - lo + 31) // 32 <= sharding.shard_words(n, world)


def _worker(rank, world, port, n_items, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    from oracle import coracle as co
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = np.load(os.path.join(HERE, "golden", "config2.npz"))
    items = np.ascontiguousarray(d["items"]).view(co.ITEM_DTYPE).reshape(-1)[:n_items]
    lo, hi = sharding.shard_bounds(n_items, world, rank)
    local = co.verify_batch(items[lo:hi], d["arena"].tobytes(), tables=[d["addrs"]], group_table=[0] * len(d["groups"]), n_threads=2)
    per = sharding.shard_words(n_items, world)
    words = np.zeros(per, dtype=np.uint32)
    words[: len(local)] = local
    full = sharding.all_gather_bitmap(torch.from_numpy(words.view(np.int32)), n_items, world).numpy().view(np.uint32)
    np.save(os.path.join(out_dir, f"bitmap_{rank}.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [3000, 1001])
def test_two_rank_gloo_bitmap_allgather_matches_unsharded(tmp_path, n_items):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n_items, str(tmp_path)), nprocs=2, join=True)
    d = np.load(os.path.join(HERE, "golden", "config2.npz"))
    want = d["bitmap"][: (n_items + 31) // 32].copy()
    if n_items & 31:
        want[-1] &= (1 << (n_items & 31)) - 1
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"bitmap_{r}.npy"))
        assert np.array_equal(got, want)     # every rank holds the complete, identical bitmap
