"""BASELINE.json configs 4 and 5 as parity-test cases (config 1-3 live in test_oracle_logic / test_gpu_verify).

config 4: ROUND-CHANGE fan-in with nested prepared certificates -- through the GPU-backed host mirror at a
          scale the oracle replays in seconds (the full 10k case is 9 GB of wire bytes; its raw nested checks are the same
          tuples the throughput bench runs).
config 5: 100k pending mixed items across 16 concurrent heights (16 validator tables), sharded 8 ways through the
          device-resident ABI; bitmap and per-(height, type) quorum results bit-exact against the oracle."""
import importlib

import numpy as np
import pytest
import torch

import ibft_b200 as ib
import workloads as wl
from oracle import coracle as co
from oracle import ibft_logic as L
from oracle import ibft_proto as ip
from test_gpu_host import gpu_ctx, oracle_backend

pytestmark = pytest.mark.gpu
host = importlib.import_module("go-ibft_b200.host")
sharding = importlib.import_module("go-ibft_b200.sharding")
enc = ip.encode_ibft_message


def signed(m, key):
    m.signature = wl.sign(key, co.keccak256(m.payload_no_sig()))
    return m


def test_config4_round_change_fan_in_three_certificates():
    n, height = 90, 1_000_000
    vs = wl.ValidatorSet(3, n)                      # unit voting power: quorum 61
    raw = bytes(range(256)) * 4
    view0, view1 = ip.View(height, 0), ip.View(height, 1)
    ph = wl.proposal_hash(raw, 0)
    pp = signed(ip.IbftMessage(view0, vs.addrs[0], b"", ip.PREPREPARE, ip.PrePrepareMessage(ip.Proposal(raw, 0), ph, None)), vs.keys[0])
    prepares = [signed(ip.IbftMessage(view0, vs.addrs[i], b"", ip.PREPARE, ip.PrepareMessage(ph)), vs.keys[i]) for i in range(1, n)]
    pcs = [ip.PreparedCertificate(pp, prepares[:60]), ip.PreparedCertificate(pp, prepares[10:75]), ip.PreparedCertificate(pp, prepares[20:85])]
    bad_pc = ip.decode_pc(ip.encode_pc(pcs[0]))
    sig = bytearray(bad_pc.prepare_messages[7].signature)
    sig[40] ^= 1
    bad_pc.prepare_messages[7].signature = bytes(sig)
    short_pc = ip.PreparedCertificate(pp, prepares[:30])          # below quorum
    rcs = []
    for i in range(n):
        pc = pcs[i % 3]
        if i % 29 == 7:
            pc = bad_pc                                            # corrupted nested signature
        if i == 11:
            pc = short_pc
        rcs.append(signed(ip.IbftMessage(view1, vs.addrs[i], b"", ip.ROUND_CHANGE, ip.RoundChangeMessage(ip.Proposal(raw, 0), pc)), vs.keys[i]))
    rcs.append(signed(ip.IbftMessage(view1, vs.addrs[5], b"", ip.ROUND_CHANGE, ip.RoundChangeMessage(None, None)), vs.keys[5]))  # overwrites node 5: no PC is valid
    proposer_of = lambda a, h, r: a == vs.addrs[0] and r == 0  # noqa: E731
    o = L.IBFT(oracle_backend({height: set(vs.addrs)}, height, proposer_of), L.ValidatorManager(lambda h: {a: 1 for a in vs.addrs}))
    o.vm.init(height)
    c = gpu_ctx(proposer_of)
    assert c.set_validators(height, vs.addrs, None) == 0
    o.state.view = view1
    c.set_state(height, 1, L.NEW_ROUND, None)
    for m in rcs:
        o.messages.add_message(m)
        c.store_add(enc(m))
    items0, calls0 = c.gpu_items_verified(), c.gpu_device_calls()
    want = o.handle_round_change_message(view1)
    got = c.handle_round_change(height, 1)
    assert want is not None and got == sorted(m.from_ for m in want.round_change_messages)
    assert len(got) == n - len([i for i in range(n) if (i % 29 == 7 or i == 11) and i != 5])
    # raw nested checks would be ~ n * 62; unique signatures verified: n RC senders + 1 PREPREPARE + the 85 distinct PREPAREs the
    # three certificates cover + 1 corrupted variant
    assert c.gpu_items_verified() - items0 == n + 1 + 85 + 1
    assert c.gpu_device_calls() - calls0 <= 3
    # a PREPREPARE for round 1 carrying this RCC: validateProposal's nested fan-out (core/ibft.go:683-788)
    rcc = ip.RoundChangeCertificate(want.round_change_messages)
    pp1 = signed(ip.IbftMessage(view1, vs.addrs[1], b"", ip.PREPREPARE,
                                ip.PrePrepareMessage(ip.Proposal(raw, 1), wl.proposal_hash(raw, 1), rcc)), vs.keys[1])
    prop1 = lambda a, h, r: (a == vs.addrs[0] and r == 0) or (a == vs.addrs[1] and r == 1)  # noqa: E731
    o2 = L.IBFT(oracle_backend({height: set(vs.addrs)}, height, prop1, node_id=vs.addrs[2]), o.vm)
    c2 = gpu_ctx(prop1, node_id=vs.addrs[2])
    assert c2.set_validators(height, vs.addrs, None) == 0
    c2.set_state(height, 1, L.NEW_ROUND, None)
    # max-round rule: the certificates are for round 0, so the expected hash is hash(raw, 0), but the proposal must carry hash(raw, 1)
    assert c2.validate_proposal(enc(pp1), height, 1) == o2.validate_proposal(pp1, view1) is True
    pp1_bad = ip.decode_ibft_message(enc(pp1))
    pp1_bad.payload.certificate.round_change_messages[3].signature = b"\x01" * 65
    assert c2.validate_proposal(enc(pp1_bad), height, 1) == o2.validate_proposal(pp1_bad, view1) is False
    c.close()
    c2.close()


def _results_equal(res, want):
    """engine group results (RESULT_DTYPE) vs the pinned table of tests/golden/make_pins.py"""
    for g in range(len(want)):
        assert (int(res[g]["n_valid"]), int(res[g]["n_distinct"]), int(res[g]["has_quorum"])) == tuple(int(x) for x in want[g, :3]), g
        assert [int(x) for x in res[g]["power"]] == [int(x) for x in want[g, 4:9]], g


@pytest.fixture(scope="module")
def big_engine():
    e = ib.Engine(device=0, max_items=1 << 18, max_payload_bytes=1 << 25, max_groups=128, max_table_slots=16, max_validators=10_000)
    yield e
    e.close()


def test_config5_full_size_100k_messages_16_heights_10k_validator_tables(big_engine):
    """BASELINE config 5 as stated: 100,000 pending messages (144,953 signature tuples), 16 concurrent heights, a 10,000-validator
    table per height, 45/45/9/1 mix, 1 % adversarial.  Bitmap and the 80 per-group quorum results bit-exact against the committed
    oracle pin -- through the host-buffer ABI in one call, and as 8 shards through the device-resident ABI (the all-gather of
    real ranks is bench.py --workload config5)."""
    w, pin = wl.load_full("config5")
    e = big_engine
    for k in range(16):
        e.set_validators(k, w["heights"][k], w["tables"][k], w["powers"])
    groups = e.groups(w["n_groups"], slot=w["group_table"])
    items, arena = w["items"], w["arena"]
    bitmap, results, _ = e.verify_batch(items, arena, groups)
    assert np.array_equal(bitmap, pin["bitmap"]), "verdict bitmap differs from the oracle pin"
    _results_equal(results, pin["results"])
    # 8 "ranks" on one device: disjoint 32-aligned shards, each with its own partial voted sets, merged as after the all-gather
    total = len(items)
    e.bind_groups(groups)
    t_items = torch.from_numpy(items.view(np.uint8).reshape(-1, 128).copy()).cuda()
    t_arena = torch.from_numpy(np.frombuffer(arena, np.uint8).copy()).cuda()
    t_bm = torch.zeros((total + 31) // 32, dtype=torch.int32, device="cuda")
    W = e.quorum_partial_words()
    t_parts = torch.zeros((8, W), dtype=torch.int32, device="cuda")
    t_res = torch.zeros(w["n_groups"] * ib.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for r in range(8):
        lo, hi = sharding.shard_bounds(total, 8, r)
        e.verify_device(t_items.data_ptr(), total, t_arena.data_ptr(), len(arena), lo, hi, t_bm.data_ptr(), 0, st)
        e.quorum_mark_device(t_items.data_ptr(), total, lo, hi, t_bm.data_ptr(), t_parts[r].data_ptr(), st)
    e.quorum_merge_device(t_parts.data_ptr(), 8, W, t_res.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(t_bm.cpu().numpy().view(np.uint32), pin["bitmap"])
    _results_equal(t_res.cpu().numpy().view(ib.RESULT_DTYPE), pin["results"])
    e.bind_groups(None)


def test_config4_full_size_10k_round_change_with_nested_certificates(big_engine):
    """BASELINE config 4 as stated, dedup mode: 10,000 ROUND_CHANGE messages of 10,000 validators, each embedding one of three
    prepared certificates (1 PREPREPARE + 6,666 PREPAREs; 100 messages carry a corrupted nested signature).  20,003 unique tuples;
    the 10,000 sender tuples are IBFT_KIND_PAYLOAD2: their signed bytes are a 1.1 KB head + a shared 909 KB certificate (9 GB of
    sponge input in all, hashed on the device from 7 resident certificate blobs).  Verdicts, per-message validity and the round's
    quorum decision bit-exact against the committed oracle pin."""
    w, pin = wl.load_full("config4_n10k")
    e = big_engine
    e.set_validators(0, w["height"], w["addrs"], None)
    bitmap, results, _ = e.verify_batch(w["items"], w["arena"], e.groups(1))
    assert np.array_equal(bitmap, pin["bitmap"]), "verdict bitmap differs from the oracle pin"
    valid, hq = wl.config4_expected(w, bitmap)
    assert np.array_equal(np.packbits(valid), pin["rc_valid"]) and int(hq) == int(pin["has_quorum"][0])
    assert int(results[0]["n_valid"]) == 20_000 and int(results[0]["n_distinct"]) == 10_000 and bool(results[0]["has_quorum"])


@pytest.mark.parametrize("raw_frames", [False, True])
def test_config4_n1000_end_to_end_through_wire_codec_and_host_mirror(raw_frames):
    """The same shape at N = 1,000 through the REAL path: wire frames -> C++ decoder -> batching store shim -> GpuVerifier (one
    de-duplicated device batch) -> handleRoundChangeMessage, against oracle/ibft_logic.py with the oracle's ecrecover
    (core/ibft.go:470-512, :1162-1231; messages/messages.go:202-245).  93 MB of ROUND_CHANGE frames, ~667k nested checks,
    1,002 + 999 unique signatures.
    raw_frames=True: every message -- the ROUND_CHANGE frames AND the messages nested in their certificates -- goes to the device as
    a span of the gossip frame it arrived in (IBFT_KIND_WIRE): zero PayloadNoSig re-marshals on the host."""
    n, height = 1000, 1_000_000
    q = 2 * n // 3 + 1
    priv = co.privkeys(3, n)
    addr = co.addresses(priv, 8)
    addrs = [bytes(a) for a in addr]
    raw = bytes(range(256)) * 4
    view0, view1 = ip.View(height, 0), ip.View(height, 1)
    ph = wl.proposal_hash(raw, 0)

    def sign_all(msgs, keys):
        dig = np.stack([np.frombuffer(co.keccak256(m.payload_no_sig()), np.uint8) for m in msgs])
        sigs = co.sign_derived_batch(keys, dig, 8)
        for m, s in zip(msgs, sigs):
            m.signature = bytes(s)
        return msgs
    pp = sign_all([ip.IbftMessage(view0, addrs[0], b"", ip.PREPREPARE, ip.PrePrepareMessage(ip.Proposal(raw, 0), ph, None))], priv[:1])[0]
    prepares = sign_all([ip.IbftMessage(view0, addrs[i], b"", ip.PREPARE, ip.PrepareMessage(ph)) for i in range(1, n)], priv[1:])
    span = q - 1
    pcs = [ip.PreparedCertificate(pp, prepares[a:a + span]) for a in (0, 150, 300)]
    bad_pc = ip.decode_pc(ip.encode_pc(pcs[0]))
    sig = bytearray(bad_pc.prepare_messages[7].signature)
    sig[40] ^= 1
    bad_pc.prepare_messages[7].signature = bytes(sig)
    short_pc = ip.PreparedCertificate(pp, prepares[: span // 2])
    rcs = []
    for i in range(n):
        pc = pcs[i % 3]
        if i % 100 == 7:
            pc = bad_pc
        if i == 11:
            pc = short_pc
        rcs.append(ip.IbftMessage(view1, addrs[i], b"", ip.ROUND_CHANGE, ip.RoundChangeMessage(ip.Proposal(raw, 0), pc)))
    # signed bytes = head || certificate: hash through the batch oracle (the Python encoder would re-walk 667k nested messages)
    blobs = {id(pc): ip.encode_pc(pc) for pc in pcs + [bad_pc, short_pc]}
    cat, offs, lens = bytearray(), [], []
    for m in rcs:
        b = blobs[id(m.payload.latest_prepared_certificate)]
        offs.append(len(cat))
        cat.extend(wl.rc_head(view1, m.from_, ip.Proposal(raw, 0), len(b)))
        cat.extend(b)
        lens.append(len(cat) - offs[-1])
    sigs = co.sign_derived_batch(priv, co.keccak256_batch(bytes(cat), offs, lens, 8), 8)
    wires = []
    for i, m in enumerate(rcs):
        m.signature = bytes(sigs[i])
        b = blobs[id(m.payload.latest_prepared_certificate)]
        # wire frame = view, from, signature (field 3), type, roundChangeData: spliced from the same pieces
        head = wl.rc_head(view1, m.from_, ip.Proposal(raw, 0), len(b))
        cut = len(ip._f_msg(1, ip.encode_view(view1)) + ip._f_bytes(2, m.from_))
        wires.append(head[:cut] + ip._f_bytes(3, m.signature) + head[cut:] + b)
    assert wires[3] == enc(rcs[3])                      # the splice IS the codec's encoding
    proposer_of = lambda a, h, r: a == addrs[0] and r == 0  # noqa: E731
    memo = {}
    inner = oracle_backend({height: set(addrs)}, height, proposer_of)

    def memo_valid(m):                                  # certificates share message objects: one ecrecover per distinct signature
        k = (m.from_, m.signature, m.type, id(m.payload) if m.type == ip.ROUND_CHANGE else 0)
        if k not in memo:
            if m.type == ip.ROUND_CHANGE:
                b = blobs[id(m.payload.latest_prepared_certificate)]
                dig = co.keccak256(wl.rc_head(view1, m.from_, ip.Proposal(raw, 0), len(b)) + b)
                memo[k] = co.ecrecover_address(dig, m.signature) == m.from_ and m.from_ in set(addrs)
            else:
                memo[k] = inner.is_valid_validator(m)
        return memo[k]
    ob = L.Backend(is_valid_validator=memo_valid, is_valid_committed_seal=inner.is_valid_committed_seal,
                   is_valid_proposal_hash=inner.is_valid_proposal_hash, is_proposer=proposer_of, id=lambda: b"")
    o = L.IBFT(ob, L.ValidatorManager(lambda h: {a: 1 for a in addrs}))
    o.vm.init(height)
    params = host.EngineParams(0, 1 << 14, 1 << 27, 32, 8, 4096, 0)   # the 1,000 sender payloads are 93 MB of signed bytes
    c = host.HostContext("gpu", {"is_proposer": proposer_of}, b"", params)
    assert c.set_validators(height, addrs, None) == 0
    c.set_wire_frames(raw_frames)
    o.state.view = view1
    c.set_state(height, 1, L.NEW_ROUND, None)
    for m, wbytes in zip(rcs, wires):
        o.messages.add_message(m)
        c.store_add(wbytes)
    items0, calls0 = c.gpu_items_verified(), c.gpu_device_calls()
    want = o.handle_round_change_message(view1)
    got = c.handle_round_change(height, 1)
    assert want is not None and got == sorted(m.from_ for m in want.round_change_messages)
    assert len(got) == n - len([i for i in range(n) if i % 100 == 7 or i == 11])
    # unique signatures: n senders + 1 PREPREPARE + the distinct PREPAREs the certificates cover (300 + span) + 1 corrupted variant
    assert c.gpu_items_verified() - items0 == n + 1 + (300 + span) + 1
    assert c.gpu_device_calls() - calls0 <= 3
    assert c.gpu_frames_handed_back() == 0              # every frame was canonical: nothing was re-marshalled on the host
    c.close()


def test_sharded_verifier_pipeline_single_rank(big_engine):
    """go-ibft_b200/sharding.py ShardedVerifier (the per-rank pipeline bench.py's strong-scaling legs run under torchrun) at
    world = 1: rebased shard, H2D, kernels, merge, D2H -- bitmap and quorum results equal to the config-5 pin; and the shard
    rebasing itself for several (world, rank) against the pin's bitmap words."""
    w, pin = wl.load_full("config5")
    e = big_engine
    for k in range(16):
        e.set_validators(k, w["heights"][k], w["tables"][k], w["powers"])
    groups = e.groups(w["n_groups"], slot=w["group_table"])
    items, arena = w["items"], np.frombuffer(w["arena"], np.uint8)
    n = len(items)
    li, la = sharding.rebase_shard(items, arena, 0, n)
    sv = sharding.ShardedVerifier(e, n, groups, 1, 0, li, la, torch.cuda.current_stream())
    res, bm = sv.run()
    assert np.array_equal(bm, pin["bitmap"])
    _results_equal(res, pin["results"])
    res2, bm2 = sv.run()                                   # the pipeline is re-runnable (buffers are reused)
    assert np.array_equal(bm2, bm) and res2.tobytes() == res.tobytes()
    # the exchange as ONE kernel over (here: its own) peer memory instead of a library collective: same answers, several rounds
    svp = sharding.ShardedVerifier(e, n, groups, 1, 0, li, la, torch.cuda.current_stream(), exchange="p2p")
    for _ in range(3):
        resp, bmp = svp.run()
        assert np.array_equal(bmp, pin["bitmap"]) and resp.tobytes() == res.tobytes()
    svp.close()
    # a rank's rebased shard verifies to exactly its slice of the bitmap
    for world, rank in ((2, 1), (8, 5), (8, 7)):
        lo, hi = sharding.shard_bounds(n, world, rank)
        li, la = sharding.rebase_shard(items, arena, lo, hi)
        assert la.size < arena.size
        bitmap, _, _ = e.verify_batch(li, la, groups)
        assert np.array_equal(bitmap, pin["bitmap"][lo // 32: (hi + 31) // 32])
    e.bind_groups(None)
