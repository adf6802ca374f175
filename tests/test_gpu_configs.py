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


def test_config5_mixed_backlog_16_heights_sharded_8_ways(engine):
    heights = [1_000_000 + k for k in range(16)]
    n_val, total = 256, 100_000
    rng = np.random.default_rng(5)
    sets = [wl.ValidatorSet(10 + k, n_val, weighted=True) for k in range(16)]
    # engine.max_table_slots = 16: one slot per height
    for k, vs in enumerate(sets):
        engine.set_validators(k, heights[k], vs.addr_array(), vs.power_array())
    raw = rng.integers(0, 256, 300, dtype=np.uint8).tobytes()
    types = rng.choice([ip.PREPARE, ip.COMMIT, ip.ROUND_CHANGE, ip.PREPREPARE], size=total, p=[0.45, 0.45, 0.09, 0.01])
    hk = rng.integers(0, 16, size=total)
    vi = rng.integers(0, n_val, size=total)
    # one signed message per (height, validator, type) is reused for repeated draws: 100k tuples from ~16k distinct signatures
    cache = {}
    items, arena = [], bytearray()
    groups = np.zeros(64, dtype=ib.GROUP_DTYPE)
    for k in range(16):
        for t in range(4):
            groups[k * 4 + t]["table_slot"] = k
    outsider = wl.privkey(9999, 0)
    for i in range(total):
        k, v, t = int(hk[i]), int(vi[i]), int(types[i])
        key = (k, v, t)
        if key not in cache:
            vs = sets[k]
            view = ip.View(heights[k], 0)
            ph = wl.proposal_hash(raw, 0)
            payload = {ip.PREPARE: ip.PrepareMessage(ph), ip.COMMIT: ip.CommitMessage(ph, b"\x07" * 65),
                       ip.ROUND_CHANGE: ip.RoundChangeMessage(None, None), ip.PREPREPARE: ip.PrePrepareMessage(ip.Proposal(raw, 0), ph, None)}[t]
            m = ip.IbftMessage(view, vs.addrs[v], b"", t, payload)
            signer_key = vs.keys[v]
            tag = (k * 131 + v * 7 + t) % 97
            if tag == 3:
                signer_key = outsider                       # From != signer
            p = m.payload_no_sig()
            sig = wl.sign(signer_key, co.keccak256(p))
            if tag == 5:
                m.view = ip.View(heights[(k + 1) % 16], 0)  # replayed on another height: payload differs -> invalid
                p = m.payload_no_sig()
            cache[key] = (sig, vs.addrs[v], p)
        sig, signer, p = cache[key]
        off = len(arena)
        arena.extend(p)
        items.append(wl.make_item(sig, signer, 1, b"", k * 4 + t, off, len(p)))
    items = np.concatenate(items)
    arena = bytes(arena)
    want = co.verify_batch(items, arena, tables=[vs.addr_array() for vs in sets], group_table=[g // 4 for g in range(64)], n_threads=8)
    assert 0.9 < sum(bin(int(w)).count("1") for w in want) / total < 0.995
    # 8 "ranks" on one device: disjoint 32-aligned shards, bitmap words assembled as the all-gather would, quorum on the whole
    engine.bind_groups(groups)
    t_items = torch.from_numpy(items.view(np.uint8).reshape(-1, 128).copy()).cuda()
    t_arena = torch.from_numpy(np.frombuffer(arena, np.uint8).copy()).cuda()
    t_bm = torch.zeros((total + 31) // 32, dtype=torch.int32, device="cuda")
    t_res = torch.zeros(64 * ib.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for r in range(8):
        lo, hi = sharding.shard_bounds(total, 8, r)
        engine.verify_device(t_items.data_ptr(), total, t_arena.data_ptr(), len(arena), lo, hi, t_bm.data_ptr(), 0, st)
    engine.quorum_reduce_device(t_items.data_ptr(), total, t_bm.data_ptr(), 64, t_res.data_ptr(), st)
    torch.cuda.synchronize()
    got = t_bm.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)
    res = t_res.cpu().numpy().view(ib.RESULT_DTYPE)
    bits = np.unpackbits(want.view(np.uint8), bitorder="little")[:total]
    for g in range(64):
        k = g // 4
        vs = sets[k]
        idx = {a: i for i, a in enumerate(vs.addrs)}
        sel = np.nonzero((items["group"] == g) & (bits == 1))[0]
        voters = {idx[bytes(items[i]["signer"])] for i in sel}
        power = sum(vs.powers[v] for v in voters)
        quorum = 2 * sum(vs.powers) // 3 + 1
        assert int(res[g]["n_valid"]) == len(sel) and int(res[g]["n_distinct"]) == len(voters)
        assert sum(int(res[g]["power"][j]) << (64 * j) for j in range(5)) == power
        assert bool(res[g]["has_quorum"]) == (power >= quorum)
    engine.bind_groups(None)
    # the same backlog through the host-buffer ABI in one call
    bm2, res2, _ = engine.verify_batch(items[:60000], arena, groups)
    assert np.array_equal(bm2, want[: 60000 // 32])
