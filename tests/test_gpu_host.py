"""GPU-backed host mirror: the reference-facing Verifier (IsValidValidator / IsValidProposalHash / IsValidCommittedSeal), the
batching store shim and the handlers, with REAL signatures, against the oracle (oracle/ibft_logic.py driven by the C
oracle's ecrecover).  Checks decisions, pruned store contents, and that batching really is one device call per handler."""
import importlib

import numpy as np
import pytest

import ibft_b200 as ib
import workloads as wl
from oracle import coracle as co
from oracle import ibft_logic as L
from oracle import ibft_proto as ip

pytestmark = pytest.mark.gpu
host = importlib.import_module("go-ibft_b200.host")
enc = ip.encode_ibft_message


def oracle_backend(vs_by_height, current_height, proposer_of, node_id=b""):
    """The embedder's Backend restated with the oracle's crypto (SURVEY.md §8c conventions)."""
    def is_valid_validator(m):
        if m.view is None or len(m.from_) != 20 or len(m.signature) != 65:
            return False
        addr = co.ecrecover_address(co.keccak256(m.payload_no_sig()), m.signature)
        return addr == m.from_ and m.from_ in vs_by_height.get(m.view.height, ())

    def is_valid_committed_seal(h, seal):
        if h is None or seal is None or len(h) != 32 or len(seal.signer) != 20 or len(seal.signature) != 65:
            return False
        addr = co.ecrecover_address(wl.seal_digest(h), seal.signature)
        return addr == seal.signer and seal.signer in vs_by_height.get(current_height, ())

    def is_valid_proposal_hash(p, h):
        return p is not None and h is not None and len(h) == 32 and wl.proposal_hash(p.raw_proposal, p.round) == h
    return L.Backend(is_valid_validator=is_valid_validator, is_valid_committed_seal=is_valid_committed_seal,
                     is_valid_proposal_hash=is_valid_proposal_hash, is_proposer=proposer_of, id=lambda: node_id)


def gpu_ctx(proposer_of, node_id=b"", flags=0):
    """flags = 1: the verifier's engine keeps a key registry (IBFT_FLAG_KEY_CACHE) -- same decisions, by construction"""
    params = host.EngineParams(0, 1 << 14, 1 << 24, 32, 8, 4096, flags)
    return host.HostContext("gpu", {"is_proposer": proposer_of}, node_id, params)


def make_round(n, height, rnd, seed=21):
    vs = wl.ValidatorSet(seed, n, weighted=True)
    raw = bytes(range(200)) * 3
    ph = wl.proposal_hash(raw, rnd)
    view = ip.View(height, rnd)

    def signed(m, key):
        m.signature = wl.sign(key, co.keccak256(m.payload_no_sig()))
        return m
    pp = signed(ip.IbftMessage(view, vs.addrs[0], b"", ip.PREPREPARE, ip.PrePrepareMessage(ip.Proposal(raw, rnd), ph, None)), vs.keys[0])
    prepares = [signed(ip.IbftMessage(view, vs.addrs[i], b"", ip.PREPARE, ip.PrepareMessage(ph)), vs.keys[i]) for i in range(1, n)]
    commits = [signed(ip.IbftMessage(view, vs.addrs[i], b"", ip.COMMIT, ip.CommitMessage(ph, wl.sign(vs.keys[i], wl.seal_digest(ph)))), vs.keys[i])
               for i in range(n)]
    return vs, raw, ph, pp, prepares, commits


@pytest.mark.parametrize("flags", [0, 1], ids=["recover", "key_registry"])
def test_commit_round_batched_and_serial_match_oracle(flags):
    n, height = 40, 1_000_000
    vs, raw, ph, pp, prepares, commits = make_round(n, height, 0)
    outsider = wl.privkey(5000, 1)
    # adversarial COMMITs: bad seal (signed by someone else), seal over another hash, outsider (non-member) with a valid seal
    commits[3].payload.committed_seal = wl.sign(vs.keys[4], wl.seal_digest(ph))
    commits[5].payload.committed_seal = wl.sign(vs.keys[5], wl.seal_digest(b"\x01" * 32))
    commits[7].payload.committed_seal = commits[7].payload.committed_seal[:64] + b"\x05"
    commits[9].payload.proposal_hash = b"\x02" * 32
    commits[11].payload.committed_seal = b"short"
    out_addr = wl.address_of(outsider)
    commits.append(ip.IbftMessage(ip.View(height, 0), out_addr, b"", ip.COMMIT, ip.CommitMessage(ph, wl.sign(outsider, wl.seal_digest(ph)))))
    proposer_of = lambda a, h, r: a == vs.addrs[0]  # noqa: E731
    for batching in (True, False):
        o = L.IBFT(oracle_backend({height: set(vs.addrs)}, height, proposer_of), L.ValidatorManager(lambda h: dict(zip(vs.addrs, vs.powers))))
        o.vm.init(height)
        c = gpu_ctx(proposer_of, flags=flags)
        c.set_batching(batching)
        assert c.set_validators(height, vs.addrs, vs.powers) == 0
        for x in (o.state,):
            x.view, x.name, x.proposal_message = ip.View(height, 0), L.COMMIT_STATE, pp
        c.set_state(height, 0, L.COMMIT_STATE, enc(pp))
        for m in commits:
            o.messages.add_message(m)
            c.store_add(enc(m))
        calls0 = c.gpu_device_calls()
        assert o.handle_commit(ip.View(height, 0)) == c.handle_commit(height, 0) is True
        assert c.store_senders(height, 0, ip.COMMIT) == sorted(m.from_ for m in o.messages.maps[ip.COMMIT][height][0].values())
        assert c.seal_count() == len(o.state.seals) == n - 5
        calls = c.gpu_device_calls() - calls0
        # batched: ONE proposal-hash launch (both sponges) + ONE verify launch for all seals; serial: one launch per seal
        assert calls == (2 if batching else 1 + len([m for m in commits if len(m.payload.committed_seal) == 65 and len(m.payload.proposal_hash) == 32 and m.payload.proposal_hash == ph]))
        c.close()


@pytest.mark.parametrize("flags", [0, 1], ids=["recover", "key_registry"])
def test_ingress_prepare_flow_and_single_calls(flags):
    n, height = 24, 77
    vs, raw, ph, pp, prepares, commits = make_round(n, height, 0, seed=22)
    proposer_of = lambda a, h, r: a == vs.addrs[0]  # noqa: E731
    o = L.IBFT(oracle_backend({height: set(vs.addrs)}, height, proposer_of), L.ValidatorManager(lambda h: dict(zip(vs.addrs, vs.powers))))
    o.vm.init(height)
    c = gpu_ctx(proposer_of, flags=flags)
    assert c.set_validators(height, vs.addrs, vs.powers) == 0
    o.state.view = ip.View(height, 0)
    c.set_state(height, 0, L.NEW_ROUND, None)
    # ingress: forged sender (From != signer), tampered payload, wrong-height table, malformed signature
    forged = ip.IbftMessage(ip.View(height, 0), vs.addrs[2], prepares[0].signature, ip.PREPARE, ip.PrepareMessage(ph))
    tampered = ip.decode_ibft_message(enc(prepares[1]))
    tampered.payload.proposal_hash = b"\x09" * 32
    other_height = ip.decode_ibft_message(enc(prepares[2]))
    other_height.view = ip.View(height + 1, 0)            # signature no longer matches AND no table for that height
    nosig = ip.IbftMessage(ip.View(height, 0), vs.addrs[3], b"", ip.PREPARE, ip.PrepareMessage(ph))
    inbound = [pp] + prepares[3:] + [forged, tampered, other_height, nosig]
    for m in inbound:
        assert c.is_valid_validator(enc(m)) == o.backend.is_valid_validator(m)
    calls0, items0 = c.gpu_device_calls(), c.gpu_items_verified()
    for m in inbound:
        o.add_message(m)
    c.add_messages([enc(m) for m in inbound])               # bulk ingress: cached verdicts, no new launches needed
    assert c.gpu_device_calls() == calls0
    assert c.num_messages(height, 0, ip.PREPARE) == o.messages.num_messages(ip.View(height, 0), ip.PREPARE) == len(prepares) - 3
    assert c.signal_count() == len(o.messages.signals)
    # fresh context: the same bulk ingress is ONE launch
    c2 = gpu_ctx(proposer_of, flags=flags)
    assert c2.set_validators(height, vs.addrs, vs.powers) == 0
    c2.set_state(height, 0, L.NEW_ROUND, None)
    c2.add_messages([enc(m) for m in inbound])
    assert c2.gpu_device_calls() == 1 and c2.gpu_items_verified() == len(inbound) - 1   # the unsigned message never reaches the device
    # proposal acceptance + prepare quorum
    got = c2.handle_preprepare(height, 0)
    want = o.handle_preprepare(ip.View(height, 0))
    assert got == want.from_ == vs.addrs[0]
    for ctx_state in (o.state,):
        ctx_state.proposal_message, ctx_state.name = pp, L.PREPARE_STATE
    c2.set_state(height, 0, L.PREPARE_STATE, enc(pp))
    assert c2.handle_prepare(height, 0) == o.handle_prepare(ip.View(height, 0)) is True
    assert c2.latest_pc_prepares() == len(o.state.latest_pc.prepare_messages)
    # verifier edge cases: nil / malformed => false, never a crash
    assert c2.is_valid_committed_seal(None, vs.addrs[0], b"\x00" * 65) is False
    assert c2.is_valid_committed_seal(ph, None) is False
    assert c2.is_valid_committed_seal(ph[:31], vs.addrs[0], b"\x00" * 65) is False
    assert c2.is_valid_proposal_hash(None, 0, ph) is False and c2.is_valid_proposal_hash(raw, 0, None) is False
    assert c2.is_valid_proposal_hash(raw, 0, ph) is True and c2.is_valid_proposal_hash(raw, 1, ph) is False
    c.close()
    c2.close()


@pytest.mark.parametrize("flags", [0, 1], ids=["recover", "key_registry"])
def test_round_change_with_nested_certificates_dedup(flags):
    """Config-4 shape at small scale: every ROUND_CHANGE embeds the same prepared certificate: the nested
    sender signatures are verified once each (dedup), in the same launch as nothing else."""
    n, height = 16, 9
    vs, raw, ph, pp0, prepares0, _ = make_round(n, height, 0, seed=23)
    quorum_prepares = prepares0[:12]                                 # proposer + validators 1..12: power 49 >= quorum 40 of 59
    pc = ip.PreparedCertificate(pp0, quorum_prepares)
    view1 = ip.View(height, 1)

    def signed(m, key):
        m.signature = wl.sign(key, co.keccak256(m.payload_no_sig()))
        return m
    rcs = [signed(ip.IbftMessage(view1, vs.addrs[i], b"", ip.ROUND_CHANGE, ip.RoundChangeMessage(ip.Proposal(raw, 0), pc)), vs.keys[i]) for i in range(n)]
    # one RC whose nested certificate carries a corrupted prepare signature
    bad_pc = ip.decode_pc(ip.encode_pc(pc))
    bad_pc.prepare_messages[2].signature = bad_pc.prepare_messages[2].signature[:10] + b"\x00" + bad_pc.prepare_messages[2].signature[11:]
    rcs[5] = signed(ip.IbftMessage(view1, vs.addrs[5], b"", ip.ROUND_CHANGE, ip.RoundChangeMessage(ip.Proposal(raw, 0), bad_pc)), vs.keys[5])
    proposer_of = lambda a, h, r: a == vs.addrs[(h + r) % n] if False else (a == vs.addrs[0] and r == 0)  # noqa: E731
    o = L.IBFT(oracle_backend({height: set(vs.addrs)}, height, proposer_of), L.ValidatorManager(lambda h: dict(zip(vs.addrs, vs.powers))))
    o.vm.init(height)
    c = gpu_ctx(proposer_of, flags=flags)
    assert c.set_validators(height, vs.addrs, vs.powers) == 0
    o.state.view = view1
    c.set_state(height, 1, L.NEW_ROUND, None)
    for m in rcs:
        o.messages.add_message(m)
        c.store_add(enc(m))
    items0, calls0 = c.gpu_items_verified(), c.gpu_device_calls()
    want = o.handle_round_change_message(view1)
    got = c.handle_round_change(height, 1)
    assert want is not None and got == sorted(m.from_ for m in want.round_change_messages) and len(got) == n - 1
    # unique signatures: n RC sender sigs + 1 PREPREPARE + |prepares| (+1 corrupted variant) -- not n * (1 + |prepares|)
    assert c.gpu_items_verified() - items0 == n + 1 + len(quorum_prepares) + 1
    assert c.gpu_device_calls() - calls0 <= 3                          # one verify launch + the proposal-hash keccaks
    c.close()


def test_quorum_from_gpu_voted_bitmap(engine):
    """core/validator_manager.go's quorum check reading the GPU bitmap: engine voted set -> host ValidatorManager."""
    d = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "config2.npz"))
    items = np.ascontiguousarray(d["items"]).view(ib.ITEM_DTYPE).reshape(-1)
    engine.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
    groups = engine.groups(len(d["groups"]))
    seal_group = list(d["groups"]).index("COMMIT_SEAL")
    sub = items[items["group"] == seal_group]
    c = host.HostContext("callback")
    addrs = [bytes(a) for a in d["addrs"]]
    assert c.set_validators(int(d["meta"][2]), addrs, [int.from_bytes(bytes(p), "big") for p in d["powers"]]) == 0
    for take in (600, 680, 1000):
        _, results, _ = engine.verify_batch(sub[:take], d["arena"], groups)
        voted = engine.voted_bitmap(seal_group, len(addrs))
        assert c.has_quorum_voted(voted) == bool(results[seal_group]["has_quorum"])
        valid_senders = [bytes(sub[i]["signer"]) for i in range(take) if (int(d["bitmap"][(2000 + i) >> 5]) >> ((2000 + i) & 31)) & 1]
        assert c.has_quorum_senders(valid_senders) == bool(results[seal_group]["has_quorum"])


def test_raw_frame_mode_same_decisions_and_fallback():
    """GpuVerifier with use_wire_frames: PREPARE/COMMIT sender signatures go to the device as raw gossip frames (no host-side
    PayloadNoSig marshal); a non-canonically encoded frame is handed back and takes the marshalled path -- same verdicts."""
    n, height = 20, 4242
    vs, raw, ph, pp, prepares, commits = make_round(n, height, 0, seed=29)
    proposer_of = lambda a, h, r: a == vs.addrs[0]  # noqa: E731
    o = L.IBFT(oracle_backend({height: set(vs.addrs)}, height, proposer_of), L.ValidatorManager(lambda h: dict(zip(vs.addrs, vs.powers))))
    o.vm.init(height)
    o.state.view = ip.View(height, 0)
    c = gpu_ctx(proposer_of)
    c.set_wire_frames(True)
    assert c.set_validators(height, vs.addrs, vs.powers) == 0
    c.set_state(height, 0, L.NEW_ROUND, None)
    frames = [enc(m) for m in prepares + commits]
    # unknown field 9 appended: decodes fine, not canonical.  protobuf-go keeps unknown fields through Clone + Marshal, so the bytes
    # a Go node hashes for this message are NOT the ones its sender signed: invalid on every node (the device hands the frame back
    # and the host re-marshals it the protobuf-go way; oracle/ibft_proto.py does the same)
    frames[3] = frames[3] + b"\x48\x01"
    frames[5] = frames[5][:-1] + bytes([frames[5][-1] ^ 1])    # corrupted payload byte: canonical frame, bad signature
    inbound = [ip.decode_ibft_message(f) for f in frames]
    for m in inbound:
        o.add_message(m)
    c.add_messages(frames)
    assert c.gpu_frames_handed_back() == 1
    for t in (ip.PREPARE, ip.COMMIT):
        assert c.store_senders(height, 0, t) == sorted(m.from_ for m in o.messages.maps[t].get(height, {}).get(0, {}).values())
    assert c.num_messages(height, 0, ip.PREPARE) == len(prepares) - 2
    # the same frame with the unknown field INSIDE the signed bytes is valid everywhere: sign the re-marshal of what will arrive
    m = ip.decode_ibft_message(enc(prepares[7]))
    m.signature = b""
    framed = enc(m) + b"\x48\x01"
    sig = wl.sign(vs.keys[8], co.keccak256(ip.payload_no_sig_from_wire(framed)))
    cut = len(ip._f_msg(1, ip.encode_view(m.view)) + ip._f_bytes(2, m.from_))
    framed = framed[:cut] + b"\x1a\x41" + sig + framed[cut:]    # signature TLV right after `from`
    assert ip.decode_ibft_message(framed).signature == sig and ip.payload_no_sig_from_wire(framed).endswith(b"\x48\x01")
    assert c.is_valid_validator(framed) is True
    assert c.gpu_frames_handed_back() == 2
    c.close()


def test_insert_proposal_seal_reverification_one_launch():
    """the >= Q committed seals handed to Backend.InsertProposal (core/backend.go:78-81) are re-verified
    in one launch (what a syncing node does for every imported block)."""
    n, height = 64, 12
    vs = wl.ValidatorSet(33, n, weighted=True)
    ph = co.keccak256(b"imported block")
    sd = wl.seal_digest(ph)
    c = gpu_ctx(lambda a, h, r: False)
    assert c.set_validators(height, vs.addrs, vs.powers) == 0
    c.set_state(height, 0)
    seals = [(vs.addrs[i], wl.sign(vs.keys[i], sd)) for i in range(n)]
    quorum = 2 * sum(vs.powers) // 3 + 1
    calls0 = c.gpu_device_calls()
    ok, nv = c.verify_committed_seals(ph, seals)
    assert ok and nv == n and c.gpu_device_calls() - calls0 == 1
    # corrupt seals until the valid ones stay just below quorum
    order = sorted(range(n), key=lambda i: -vs.powers[i])
    kept, power = [], 0
    for i in order:
        if power + vs.powers[i] < quorum:
            kept.append(i)
            power += vs.powers[i]
    bad = [(vs.addrs[i], wl.sign(vs.keys[i], sd)) if i in kept else (vs.addrs[i], wl.sign(vs.keys[(i + 1) % n], sd)) for i in range(n)]
    ok, nv = c.verify_committed_seals(ph, bad)
    assert not ok and nv == len(kept)
    extra = [i for i in range(n) if i not in kept][0]
    ok, nv = c.verify_committed_seals(ph, bad + [seals[extra]])
    assert nv == len(kept) + 1 and ok == (power + vs.powers[extra] >= quorum)
    c.close()
