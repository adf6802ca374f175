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


def _worker_config5(rank, world, port, out_dir):
    """Strong-scaling pipeline of go-ibft_b200/sharding.py on CPU: every rank holds ONLY its rebased shard of the config-5 backlog,
    verifies it (oracle verdicts stand in for the kernels), marks its votes, and ONE all-gather carries (bitmap words | partial voted
    sets | valid counts); every rank then merges and reduces -- same layout as ShardedVerifier / ibft_quorum_merge_device."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import workloads as wl
    from oracle import coracle as co
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, pin = wl.load_full("config5")
    items, arena = w["items"], np.frombuffer(w["arena"], np.uint8)
    n = len(items)
    lo, hi = sharding.shard_bounds(n, world, rank)
    local, local_arena = sharding.rebase_shard(items, arena, lo, hi)
    assert local_arena.size < arena.size * 0.8                       # the rank holds its own payload bytes only (the seals at the tail carry none)
    bm = co.verify_batch(local, local_arena.tobytes(), tables=w["tables"], group_table=w["group_table"], n_threads=4)
    per = sharding.shard_words(n, world)
    n_val = len(w["tables"][0])
    vw = (n_val + 31) // 32
    ng = w["n_groups"]
    part = np.zeros(per + ng * vw + ng, dtype=np.uint32)            # bitmap words | voted sets | valid counts
    part[: len(bm)] = bm
    bits = np.unpackbits(bm.view(np.uint8), bitorder="little")[: hi - lo]
    index = [{bytes(a): i for i, a in enumerate(t)} for t in w["tables"]]
    for i in np.nonzero(bits)[0]:
        g = int(local["group"][i])
        v = index[w["group_table"][g]][bytes(local["signer"][i])]
        part[per + g * vw + (v >> 5)] |= np.uint32(1 << (v & 31))
        part[per + ng * vw + g] += 1
    gathered = torch.empty(world * len(part), dtype=torch.int32)
    dist.all_gather_into_tensor(gathered, torch.from_numpy(part.view(np.int32)))
    parts = gathered.numpy().view(np.uint32).reshape(world, -1)
    bitmap = parts[:, :per].reshape(-1)[: (n + 31) // 32]
    voted = np.bitwise_or.reduce(parts[:, per: per + ng * vw], axis=0).reshape(ng, vw)
    counts = parts[:, per + ng * vw:].sum(axis=0)
    pw = [int.from_bytes(bytes(p), "big") for p in w["powers"]]
    quorum = 2 * sum(pw) // 3 + 1
    res = np.zeros((ng, 3), dtype=np.uint64)
    for g in range(ng):
        vs = np.nonzero(np.unpackbits(voted[g].view(np.uint8), bitorder="little")[:n_val])[0]
        res[g] = (counts[g], len(vs), int(sum(pw[v] for v in vs) >= quorum))
    np.savez(os.path.join(out_dir, f"c5_{rank}.npz"), bitmap=bitmap, res=res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_config5_strong_scaling_pipeline(tmp_path):
    import sys
    sys.path.insert(0, HERE)
    import workloads as wl
    w, pin = wl.load_full("config5")          # (also warms the cache the workers read)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker_config5, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"c5_{r}.npz"))
        assert np.array_equal(got["bitmap"], pin["bitmap"])
        assert np.array_equal(got["res"], pin["results"][:, :3])


class _FakeExchangeEngine:
    """stands in for Engine.exchange_open / _close / _free: 'maps' a handle to an address derived from it, or refuses"""

    def __init__(self, refuse: bool):
        self.refuse, self.closed, self.freed = refuse, [], []

    def exchange_open(self, handle: bytes) -> int:
        if self.refuse:
            raise RuntimeError("cudaIpcOpenMemHandle: peer access is not supported between these two devices")
        return 0x1000 + handle[0]

    def exchange_close(self, p: int):
        self.closed.append(p)

    def exchange_free(self, p: int):
        self.freed.append(p)


def _worker_peers(rank, world, port, out_dir, failing_rank):
    import json
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeExchangeEngine(refuse=(rank == failing_rank))
    own_ptr, handle = 0xA000 + rank, bytes([rank]) * 64
    try:
        ptrs = sharding.open_peer_buffers(eng, world, rank, own_ptr, handle)
        res = {"ptrs": ptrs}
    except RuntimeError as ex:
        res = {"error": str(ex)}
    res.update(closed=eng.closed, freed=eng.freed)
    dist.barrier()          # every rank is still in step after the collective setup, whatever its outcome
    with open(os.path.join(out_dir, f"peers_{rank}.json"), "w") as f:
        json.dump(res, f)
    dist.destroy_process_group()


@pytest.mark.parametrize("failing_rank", [-1, 1])
def test_peer_buffer_setup_is_all_or_nothing(tmp_path, failing_rank):
    """open_peer_buffers (the handle exchange of ShardedVerifier's peer-memory exchange): all ranks get the full address list, or --
    when ONE rank cannot map a peer -- all ranks raise, release what they mapped and stay in step for the next collective."""
    import json
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker_peers, args=(2, port, str(tmp_path), failing_rank), nprocs=2, join=True)
    out = [json.load(open(os.path.join(str(tmp_path), f"peers_{r}.json"))) for r in range(2)]
    if failing_rank < 0:
        assert out[0]["ptrs"] == [0xA000, 0x1001] and out[1]["ptrs"] == [0x1000, 0xA001]
        assert not any(o["closed"] or o["freed"] for o in out)
    else:
        assert all("rank 1" in o["error"] for o in out)
        assert out[0]["closed"] == [0x1001] and out[0]["freed"] == [0xA000]      # rank 0 had mapped its peer: unmapped again
        assert out[1]["closed"] == [] and out[1]["freed"] == [0xA001]
