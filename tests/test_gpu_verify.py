"""Parity tests proper: the CUDA path (through the C ABI) against the oracle and the committed golden fixtures."""
import os
import random

import numpy as np
import pytest

import ibft_b200 as ib
import workloads as wl
from oracle import coracle as co
from oracle import secp256k1 as ec

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def load_fixture(name):
    d = np.load(os.path.join(HERE, "golden", name))
    items = np.ascontiguousarray(d["items"]).view(ib.ITEM_DTYPE).reshape(-1)
    return d, items


def groups_for(eng, n_groups, slot=0):
    """group descriptors for `slot`, carrying the height of the table resident there (ibft_group_desc.height)"""
    return eng.groups(n_groups, slot)


def expected_groups(items, bitmap, addrs, powers):
    """Python big-int restatement of HasQuorum over the verdict bitmap (core/validator_manager.go:77-96, :130-135)."""
    index = {bytes(a): i for i, a in enumerate(addrs)}
    pw = [int.from_bytes(bytes(p), "big") for p in powers]
    quorum = 2 * sum(pw) // 3 + 1
    out = {}
    for i in range(len(items)):
        if (int(bitmap[i >> 5]) >> (i & 31)) & 1:
            g = int(items[i]["group"])
            s = out.setdefault(g, [0, set()])
            s[0] += 1
            s[1].add(index[bytes(items[i]["signer"])])
    return {g: (nv, len(vs), sum(pw[v] for v in vs), sum(pw[v] for v in vs) >= quorum) for g, (nv, vs) in out.items()}, quorum


@pytest.mark.parametrize("name", ["config2.npz", "config3.npz"])
def test_fixture_bitmap_and_quorum_bit_exact(engine, name):
    d, items = load_fixture(name)
    engine.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
    ng = len(d["groups"])
    bitmap, results, recovered = engine.verify_batch(items, d["arena"], groups_for(engine, ng), want_recovered=True)
    assert np.array_equal(bitmap, d["bitmap"]), "verdict bitmap differs from the oracle's golden bitmap"
    # recompute the oracle live as well (guards against a stale fixture)
    gt = [0] * ng
    assert np.array_equal(bitmap, co.verify_batch(items, d["arena"].tobytes(), tables=[d["addrs"]], group_table=gt, n_threads=8))
    exp, quorum = expected_groups(items, bitmap, d["addrs"], d["powers"])
    q, h, n = engine.get_quorum(0)
    assert (q, h, n) == (quorum, int(d["meta"][2]), len(d["addrs"]))
    for g in range(ng):
        nv, nd, power, hq = exp.get(g, (0, 0, 0, False))
        r = results[g]
        assert (int(r["n_valid"]), int(r["n_distinct"]), bool(r["has_quorum"])) == (nv, nd, hq)
        assert sum(int(r["power"][k]) << (64 * k) for k in range(5)) == power
        voted = engine.voted_bitmap(g, len(d["addrs"]))
        assert sum(bin(int(w)).count("1") for w in voted) == nd
    # recovered addresses: equal to the expected signer exactly where the signature itself verifies
    for i in random.Random(1).sample(range(len(items)), 300):
        it = items[i]
        z = (d["arena"][int(it["payload_off"]):int(it["payload_off"]) + int(it["payload_len"])].tobytes() if it["kind"] == 1 else None)
        dig = co.keccak256(z) if z is not None else wl.seal_digest(bytes(it["digest"])) if it["kind"] == 2 else bytes(it["digest"])
        want = co.ecrecover_address(dig, bytes(it["r"]) + bytes(it["s"]) + bytes([int(it["v"])]))
        assert bytes(recovered[i]) == (want or bytes(20))


def test_quorum_threshold_edge(engine):
    """Exactly quorum-1 / quorum valid seals of a 100-validator set; duplicates do not add power."""
    vs = wl.ValidatorSet(7, 100, weighted=True)
    ph = os.urandom(32)
    sd = wl.seal_digest(ph)
    total = sum(vs.powers)
    quorum = 2 * total // 3 + 1
    engine.set_validators(1, 5, vs.addr_array(), vs.power_array())
    order = list(range(100))
    acc, k = 0, 0
    while acc < quorum:
        acc += vs.powers[order[k]]
        k += 1
    for take, want in ((k - 1, False), (k, True)):
        its = [wl.make_item(wl.sign(vs.keys[i], sd), vs.addrs[i], 2, ph, 0) for i in order[:take]]
        its += [its[0].copy(), its[0].copy()]  # duplicate sender: counted once (HasQuorum works on an address set)
        bitmap, results, _ = engine.verify_batch(np.concatenate(its), b"", groups_for(engine, 1, slot=1))
        assert int(results[0]["n_valid"]) == take + 2 and int(results[0]["n_distinct"]) == take
        assert bool(results[0]["has_quorum"]) == want
        assert sum(int(results[0]["power"][j]) << (64 * j) for j in range(5)) == sum(vs.powers[i] for i in order[:take])


def test_huge_voting_powers(engine):
    """Arbitrary-precision powers (the reference uses big.Int): 2^200-scale values, exact threshold."""
    vs = wl.ValidatorSet(8, 4)
    powers = [2**200 + 5, 2**200, 2**199, 1]
    pw = np.frombuffer(b"".join(p.to_bytes(32, "big") for p in powers), np.uint8).reshape(4, 32)
    engine.set_validators(2, 9, vs.addr_array(), pw)
    q, _, _ = engine.get_quorum(2)
    assert q == 2 * sum(powers) // 3 + 1
    ph = os.urandom(32)
    sd = wl.seal_digest(ph)
    for subset in ([0, 1], [0, 2, 3], [0, 1, 2], [1, 2, 3]):
        its = np.concatenate([wl.make_item(wl.sign(vs.keys[i], sd), vs.addrs[i], 2, ph, 0) for i in subset])
        _, results, _ = engine.verify_batch(its, b"", groups_for(engine, 1, slot=2))
        s = sum(powers[i] for i in subset)
        assert sum(int(results[0]["power"][j]) << (64 * j) for j in range(5)) == s
        assert bool(results[0]["has_quorum"]) == (s >= q)


def test_zero_total_power_rejected(engine):
    vs = wl.ValidatorSet(9, 3)
    with pytest.raises(ib.EngineError) as ei:
        engine.set_validators(3, 1, vs.addr_array(), np.zeros((3, 32), np.uint8))
    assert ei.value.code == 5  # errVotingPowerNotCorrect, core/validator_manager.go:66-68


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 127, 128, 129, 1000])
def test_ragged_sizes(engine, n):
    d, items = load_fixture("config2.npz")
    engine.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
    sub = items[2000 - n // 2: 2000 - n // 2 + n].copy()  # straddles COMMIT sender sigs / seals
    bitmap, _, _ = engine.verify_batch(sub, d["arena"], groups_for(engine, len(d["groups"])))
    want = co.verify_batch(sub, d["arena"].tobytes(), tables=[d["addrs"]], group_table=[0] * len(d["groups"]), n_threads=4)
    assert np.array_equal(bitmap, want)


def test_malformed_items_are_false_never_crash(engine):
    vs = wl.ValidatorSet(10, 4)
    engine.set_validators(4, 1, vs.addr_array(), None)
    dig = os.urandom(32)
    good = wl.make_item(wl.sign(vs.keys[0], dig), vs.addrs[0], 0, dig, 0)
    bad_kind = good.copy(); bad_kind["kind"] = 77
    invalid = good.copy(); invalid["kind"] = 255                      # host-flagged: nil seal / wrong signature length
    oob = good.copy(); oob["kind"] = 1; oob["payload_off"] = 2**31; oob["payload_len"] = 100
    bad_group = good.copy(); bad_group["group"] = 9
    short_sig = wl.make_item(b"\x01" * 64, vs.addrs[0], 0, dig, 0)     # len != 65 -> KIND_INVALID
    its = np.concatenate([good, bad_kind, invalid, oob, bad_group, short_sig, good])
    bitmap, results, _ = engine.verify_batch(its, b"\x00" * 16, groups_for(engine, 1, slot=4))
    assert int(bitmap[0]) == 0b1000001
    assert int(results[0]["n_valid"]) == 2 and int(results[0]["n_distinct"]) == 1
    # a group that references an unset table slot is an error, not a verdict
    e2 = ib.Engine(device=0, max_items=64, max_payload_bytes=1024, max_groups=4, max_table_slots=4, max_validators=16)
    with pytest.raises(ib.EngineError) as ei:
        e2.verify_batch(its, b"", groups_for(e2, 1, slot=3))
    assert ei.value.code == 6
    e2.close()
    # no table: membership skipped, pure recover+compare
    g = groups_for(engine, 1, slot=ib.NO_TABLE)
    outsider = wl.privkey(999, 0)
    it2 = wl.make_item(wl.sign(outsider, dig), wl.address_of(outsider), 0, dig, 0)
    bitmap, results, _ = engine.verify_batch(it2, b"", g)
    assert int(bitmap[0]) == 1 and int(results[0]["has_quorum"]) == 0
    # same item against a table it is not a member of
    bitmap, _, _ = engine.verify_batch(it2, b"", groups_for(engine, 1, slot=4))
    assert int(bitmap[0]) == 0


def test_multiblock_payloads_and_keccak_batch(engine):
    """PREPREPARE / ROUND_CHANGE payloads span many 136-byte blocks (up to 909 KB)."""
    rnd = random.Random(3)
    vs = wl.ValidatorSet(11, 3)
    engine.set_validators(5, 1, vs.addr_array(), None)
    sizes = [0, 1, 135, 136, 137, 271, 272, 273, 1095, 5000, 70000]
    arena = bytearray()
    its = []
    msgs = []
    for k, sz in enumerate(sizes):
        payload = bytes(rnd.getrandbits(8) for _ in range(sz))
        msgs.append(payload)
        key = vs.keys[k % 3]
        its.append(wl.make_item(wl.sign(key, co.keccak256(payload)), vs.addrs[k % 3], 1, b"", 0, len(arena), sz))
        arena.extend(payload)
    its = np.concatenate(its)
    bitmap, _, _ = engine.verify_batch(its, bytes(arena), groups_for(engine, 1, slot=5))
    assert int(bitmap[0]) == (1 << len(sizes)) - 1
    assert engine.keccak256_batch(msgs) == [co.keccak256(m) for m in msgs]
    assert engine.keccak256_batch([]) == []


def test_async_submit_poll_wait(engine):
    d, items = load_fixture("config2.npz")
    engine.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
    groups = groups_for(engine, len(d["groups"]))
    bitmap = np.zeros((len(items) + 31) // 32, np.uint32)
    results = np.zeros(len(groups), ib.RESULT_DTYPE)
    arena = np.ascontiguousarray(d["arena"])
    engine.verify_submit(items, arena, groups, bitmap, results)
    with pytest.raises(ib.EngineError):
        engine.verify_submit(items, arena, groups, bitmap, results)  # one call in flight per engine
    while not engine.poll():
        pass
    engine.wait()
    assert np.array_equal(bitmap, d["bitmap"])
    assert engine.poll()


def test_device_resident_and_sharded_path(engine):
    """Device-pointer ABI with torch-owned memory: two 'ranks' verify disjoint 32-aligned shards, the bitmap words are
    concatenated (what the NCCL all-gather does) and the quorum reduce runs on the complete bitmap."""
    import torch
    d, items = load_fixture("config2.npz")
    engine.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
    n = len(items)
    groups = groups_for(engine, len(d["groups"]))
    engine.bind_groups(groups)
    t_items = torch.from_numpy(items.view(np.uint8).reshape(-1, 128).copy()).cuda()
    t_arena = torch.from_numpy(np.ascontiguousarray(d["arena"]).copy()).cuda()
    words = (n + 31) // 32
    t_bitmap = torch.zeros(words, dtype=torch.int32, device="cuda")
    t_results = torch.zeros(len(groups) * ib.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    mid = (n // 2) & ~31
    before = engine.launch_count()
    engine.verify_device(t_items.data_ptr(), n, t_arena.data_ptr(), t_arena.numel(), 0, mid, t_bitmap.data_ptr(), 0, st)
    engine.verify_device(t_items.data_ptr(), n, t_arena.data_ptr(), t_arena.numel(), mid, n, t_bitmap.data_ptr(), 0, st)
    engine.quorum_reduce_device(t_items.data_ptr(), n, t_bitmap.data_ptr(), len(groups), t_results.data_ptr(), st)
    torch.cuda.synchronize()
    assert engine.launch_count() - before == 4
    got = t_bitmap.cpu().numpy().view(np.uint32)
    if n & 31:
        got[-1] &= (1 << (n & 31)) - 1
    assert np.array_equal(got, d["bitmap"])
    res = t_results.cpu().numpy().view(ib.RESULT_DTYPE)
    exp, _ = expected_groups(items, got, d["addrs"], d["powers"])
    for g in range(len(groups)):
        assert int(res[g]["n_distinct"]) == exp[g][1] and bool(res[g]["has_quorum"]) == exp[g][3]
    with pytest.raises(ib.EngineError):
        engine.verify_device(t_items.data_ptr(), n, 0, 0, 5, n, t_bitmap.data_ptr())  # unaligned shard
    engine.bind_groups(None)


def test_replicated_large_batch_properties(engine):
    """BASELINE-size property check (oracle too slow to replay 2^16 items every run): a replicated batch must give the
    replicated bitmap, and permuting items permutes verdict bits."""
    d, items = load_fixture("config3.npz")
    engine.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
    reps = 3
    big = np.tile(items, reps)
    perm = np.random.default_rng(0).permutation(len(big))
    groups = groups_for(engine, len(d["groups"]))
    bm, results, _ = engine.verify_batch(big[perm], d["arena"], groups)
    bits = np.unpackbits(bm.view(np.uint8), bitorder="little")[: len(big)]
    want_bits = np.tile(np.unpackbits(d["bitmap"].view(np.uint8), bitorder="little")[: len(items)], reps)[perm]
    assert np.array_equal(bits, want_bits)
    # replication adds no voting power: distinct senders, not messages, are counted
    exp, _ = expected_groups(items, d["bitmap"], d["addrs"], d["powers"])
    for g in range(len(groups)):
        assert int(results[g]["n_distinct"]) == exp[g][1] and int(results[g]["n_valid"]) == reps * exp[g][0]


def test_shard_local_quorum_mark_and_merge(engine):
    """Multi-GPU form of the quorum reduction on one device: two 'ranks' mark only their own shard into partial buffers, the
    partials are merged (OR / sum) and reduced -- identical to the single-pass reduction over the whole bitmap."""
    import torch
    d, items = load_fixture("config3.npz")
    engine.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
    n = len(items)
    groups = groups_for(engine, len(d["groups"]))
    engine.bind_groups(groups)
    t_items = torch.from_numpy(items.view(np.uint8).reshape(-1, 128).copy()).cuda()
    t_arena = torch.from_numpy(np.ascontiguousarray(d["arena"]).copy()).cuda()
    t_bm = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    engine.verify_device(t_items.data_ptr(), n, t_arena.data_ptr(), t_arena.numel(), 0, n, t_bm.data_ptr(), 0, st)
    W = engine.quorum_partial_words()
    assert W == len(groups) * ((len(d["addrs"]) + 31) // 32) + len(groups)
    parts = torch.full((3, W + 5), -1, dtype=torch.int32, device="cuda")     # stride larger than W, garbage-filled
    bounds = [(0, 6400), (6400, 13312), (13312, n)]
    for r, (lo, hi) in enumerate(bounds):
        engine.quorum_mark_device(t_items.data_ptr(), n, lo, hi, t_bm.data_ptr(), parts[r].data_ptr(), st)
    res_a = torch.zeros(len(groups) * ib.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    res_b = torch.zeros_like(res_a)
    engine.quorum_merge_device(parts.data_ptr(), 3, W + 5, res_a.data_ptr(), st)
    engine.quorum_reduce_device(t_items.data_ptr(), n, t_bm.data_ptr(), len(groups), res_b.data_ptr(), st)
    torch.cuda.synchronize()
    a, b = res_a.cpu().numpy().view(ib.RESULT_DTYPE), res_b.cpu().numpy().view(ib.RESULT_DTYPE)
    assert a.tobytes() == b.tobytes() and int(a[0]["n_valid"]) > 9000
    engine.bind_groups(None)


def test_no_groups_and_capacity_errors(engine):
    """groups=NULL: pure recover + signer compare (no membership, no quorum); over-capacity batches are refused, not truncated."""
    d, items = load_fixture("config2.npz")
    sub = items[:200].copy()
    bitmap, results, _ = engine.verify_batch(sub, d["arena"], None)
    assert results is None
    want = co.verify_batch(sub, d["arena"].tobytes())          # oracle without tables
    assert np.array_equal(bitmap, want)
    small = ib.Engine(device=0, max_items=64, max_payload_bytes=256, max_groups=2, max_table_slots=2, max_validators=8)
    with pytest.raises(ib.EngineError) as ei:
        small.verify_batch(items[:100], b"", None)
    assert ei.value.code == 4                                  # IBFT_ERR_CAPACITY
    with pytest.raises(ib.EngineError) as ei:
        small.verify_batch(items[:10], bytes(1000), None)
    assert ei.value.code == 4
    with pytest.raises(ib.EngineError) as ei:
        small.set_validators(0, 1, np.zeros((9, 20), np.uint8), None)
    assert ei.value.code == 4
    small.close()


def test_random_tuples_recovered_addresses_match_oracle(engine):
    """Fuzz: 20k random (r, s, v, digest) tuples -- half of all r are not abscissas, v is often invalid, some r/s are out of
    range -- the recovered address (or its absence) must equal the oracle's for every one of them."""
    rng = np.random.default_rng(123)
    n = 20000
    items = np.zeros(n, dtype=ib.ITEM_DTYPE)
    items["r"] = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    items["s"] = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    items["digest"] = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    items["v"] = rng.choice([0, 1, 0, 1, 2, 27, 28, 255], size=n)
    items["r"][::97] = 0
    items["s"][::89] = 0
    items["r"][5::101] = 0xFF                                  # >= n
    items["digest"][::53] = 0                                   # z = 0
    _, _, recovered = engine.verify_batch(items, b"", None, want_recovered=True)
    import ctypes
    lib = co.lib()
    out = (ctypes.c_uint8 * 20)()
    n_rec = 0
    for i in range(n):
        ok = lib.oracle_ecrecover_address(items["digest"][i].ctypes.data_as(ctypes.c_void_p), items["r"][i].ctypes.data_as(ctypes.c_void_p),
                                          items["s"][i].ctypes.data_as(ctypes.c_void_p), ctypes.c_uint8(int(items["v"][i])), out)
        want = bytes(out) if ok else bytes(20)
        n_rec += bool(ok)
        assert bytes(recovered[i]) == want, i
    assert 2000 < n_rec < 8000


def test_auto_path_selection_boundaries():
    """AUTO picks the recover kernel by batch size (qsplit <= SMs*48 < split <= SMs*192 < thread) and, for a mid-size host
    batch with quorum results, uploads the round in four pieces on four streams with the votes recorded by the recover kernels.
    Every boundary size must give the golden bitmap and the same quorum sums as the reference-order computation."""
    d, items = load_fixture("config3.npz")
    eng = ib.Engine(device=0, max_items=1 << 16, max_payload_bytes=1 << 22, max_groups=8, max_table_slots=2, max_validators=16384)
    try:
        sms = eng.device_info()["sm_count"]
        eng.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
        groups = groups_for(eng, len(d["groups"]))
        gold = np.unpackbits(d["bitmap"].view(np.uint8), bitorder="little")[: len(items)]
        big = np.tile(items, 3)
        gold_big = np.tile(gold, 3)
        sizes = sorted({1, 23, 24, 25, sms * 24, sms * 24 + 1, sms * 48, sms * 48 + 1, 10000, sms * 96 - 5, sms * 96, sms * 96 + 1,
                        sms * 192, sms * 192 + 1, 40000})
        for n in sizes:
            sub = np.ascontiguousarray(big[:n])
            bm, results, rec = eng.verify_batch(sub, d["arena"], groups, want_recovered=True)
            bits = np.unpackbits(bm.view(np.uint8), bitorder="little")[:n]
            assert np.array_equal(bits, gold_big[:n]), n
            exp, _ = expected_groups(sub, bm, d["addrs"], d["powers"])
            for g in range(len(groups)):
                nv, nd, power, hq = exp.get(g, (0, 0, 0, False))
                r = results[g]
                assert (int(r["n_valid"]), int(r["n_distinct"]), bool(r["has_quorum"])) == (nv, nd, hq), (n, g)
                assert sum(int(r["power"][k]) << (64 * k) for k in range(5)) == power, (n, g)
            # the same batch without quorum results and from pinned memory takes the other upload branches
            import torch
            pinned = torch.from_numpy(sub.view(np.uint8)).pin_memory().numpy().view(ib.ITEM_DTYPE).reshape(-1)
            bm2, res2, _ = eng.verify_batch(pinned, d["arena"], None)
            want2 = bits & np.array([1] * n, dtype=np.uint8)  # without groups there is no membership check: superset of `bits`
            bits2 = np.unpackbits(bm2.view(np.uint8), bitorder="little")[:n]
            assert res2 is None and np.all(bits2 >= want2), n
    finally:
        eng.close()
