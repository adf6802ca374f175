"""Differential tests: the C++ host-side mirror (go-ibft_b200/host, through its C API) against the restatement of the
reference (oracle/ibft_logic.py) on identical wire bytes and identical verifier closures.  CPU-only (callback verifier =
the reference's mockBackend).  The reference's own decision tables are replayed through BOTH implementations."""
import importlib
import json
import os
import random

import pytest

from oracle import ibft_logic as L
from oracle import ibft_proto as ip
from test_oracle_logic import (HASH, QUORUM_CASES, SENDER, append_hash, gen_messages, gen_unique, gen_with_sender, pc_all_same,
                               set_round, vp_for_cnt)

host = importlib.import_module("go-ibft_b200.host")
HERE = os.path.dirname(os.path.abspath(__file__))
enc = ip.encode_ibft_message


def test_codec_matches_reference_descriptor_golden():
    for c in json.load(open(os.path.join(HERE, "golden", "proto_wire.json")))["cases"]:
        w = bytes.fromhex(c["wire"])
        assert host.reencode(w, True).hex() == c["wire"], c["name"]
        assert host.reencode(w, False).hex() == c["payload_no_sig"], c["name"]   # PayloadNoSig, helper.go:13-27
    assert host.reencode(bytes.fromhex("0a05"), True) is None                        # truncated -> decode error, no crash


def test_cpp_remarshal_and_merged_model_match_protobuf_on_non_canonical_frames():
    """C++ twin of oracle remarshal(): protobuf-go's Unmarshal + Marshal restated (unknown fields kept, duplicates merged, a later
    oneof member replaces an earlier one), against google.protobuf's output for the same frames; and the C++ MODEL of such a frame
    equals the oracle's merged model."""
    nc = json.load(open(os.path.join(HERE, "golden", "proto_wire.json")))["noncanonical"]
    for c in nc:
        wire = bytes.fromhex(c["wire"])
        if not c["parses"]:
            assert host.remarshal(wire) is None
            continue
        assert host.remarshal(wire).hex() == c["remarshal"], c["name"]
        assert host.remarshal(wire, with_signature=False).hex() == c["payload_no_sig"], c["name"]
        assert host.reencode(wire) == ip.encode_ibft_message(ip.decode_ibft_message(wire)), c["name"]


@pytest.mark.parametrize("powers,signers,want", QUORUM_CASES)
def test_quorum_table_cpp(powers, signers, want):
    """core/validator_manager_test.go:18-187 through the C++ ValidatorManager."""
    ctx = host.HostContext("callback")
    names = sorted(powers)
    assert ctx.set_validators(0, [n.encode() for n in names], [powers[n] for n in names]) == 0
    assert ctx.has_quorum_senders([s.encode() for s in signers]) == want
    # the same decision from a voted-set bitmap (the form the GPU produces)
    word = sum(1 << names.index(s) for s in signers)
    assert ctx.has_quorum_voted([word]) == want


def test_zero_total_power_and_uninitialised_cpp():
    ctx = host.HostContext("callback")
    assert ctx.has_quorum_senders([b"A"]) is False                         # validator_manager.go:82-84
    assert ctx.set_validators(0, [b"A", b"B"], [0, 0]) == 5                # errVotingPowerNotCorrect
    assert ctx.set_validators(0, [b"A"], [2**255]) == 0 and ctx.has_quorum_senders([b"A"])


class Dual:
    """Runs the oracle IBFT and the C++ IBFT side by side with the same closures (given on decoded oracle messages)."""

    def __init__(self, n=4, node_id=b"", is_valid_validator=None, is_proposer=None, is_valid_proposal_hash=None,
                 is_valid_committed_seal=None, is_valid_proposal=None, init=True, powers=None, extra_validators=()):
        fns_o, fns_c = {}, {}
        if is_valid_validator:
            fns_o["is_valid_validator"] = is_valid_validator
            fns_c["is_valid_validator"] = lambda wire: is_valid_validator(ip.decode_ibft_message(wire))
        if is_proposer:
            fns_o["is_proposer"] = fns_c["is_proposer"] = is_proposer
        if is_valid_proposal:
            fns_o["is_valid_proposal"] = fns_c["is_valid_proposal"] = is_valid_proposal
        if is_valid_proposal_hash:
            fns_o["is_valid_proposal_hash"] = is_valid_proposal_hash
            fns_c["is_valid_proposal_hash"] = lambda pw, h: is_valid_proposal_hash(ip.decode_proposal(pw) if pw is not None else None, h)
        if is_valid_committed_seal:
            fns_o["is_valid_committed_seal"] = is_valid_committed_seal
            fns_c["is_valid_committed_seal"] = lambda h, s: is_valid_committed_seal(h, L.CommittedSeal(*s) if s else None)
        fns_o["id"] = lambda: node_id
        vp = dict(vp_for_cnt(n)(0))
        for a in extra_validators:
            vp[a] = 1
        if powers:
            vp = powers
        self.vm = L.ValidatorManager(lambda h: vp)
        self.o = L.IBFT(L.Backend(**fns_o), self.vm)
        self.c = host.HostContext("callback", fns_c, node_id)
        if init:
            self.vm.init(0)
            names = sorted(vp)
            assert self.c.set_validators(0, names, [vp[k] for k in names]) == 0

    def set_state(self, h, r, name=L.NEW_ROUND, proposal=None):
        self.o.state.view = ip.View(h, r)
        self.o.state.name = name
        self.o.state.proposal_message = proposal
        self.c.set_state(h, r, name, enc(proposal) if proposal is not None else None)

    def store_add(self, m):
        self.o.messages.add_message(m)
        self.c.store_add(enc(m))

    def add_message(self, m):
        self.o.add_message(m)
        self.c.add_message(enc(m))

    def check_store(self, h, r, t):
        want = sorted(m.from_ for m in self.o.messages.maps[t].get(h, {}).get(r, {}).values())
        assert self.c.store_senders(h, r, t) == want
        assert self.c.num_messages(h, r, t) == len(want)

    def handle_commit(self, h, r):
        a, b = self.o.handle_commit(ip.View(h, r)), self.c.handle_commit(h, r)
        assert a == b
        assert self.c.state_name() == self.o.state.name and self.c.seal_count() == len(self.o.state.seals)
        self.check_store(h, r, ip.COMMIT)
        return a

    def handle_prepare(self, h, r):
        a, b = self.o.handle_prepare(ip.View(h, r)), self.c.handle_prepare(h, r)
        assert a == b and self.c.state_name() == self.o.state.name
        if a:
            assert self.c.latest_pc_prepares() == len(self.o.state.latest_pc.prepare_messages)
        self.check_store(h, r, ip.PREPARE)
        return a

    def handle_preprepare(self, h, r):
        a = self.o.handle_preprepare(ip.View(h, r))
        b = self.c.handle_preprepare(h, r)
        # Go map order is unspecified; with several valid proposals the reference returns "one of them": compare validity sets
        self.check_store(h, r, ip.PREPREPARE)
        assert (a is None) == (b is None)
        return a

    def handle_round_change(self, h, r):
        a = self.o.handle_round_change_message(ip.View(h, r))
        b = self.c.handle_round_change(h, r)
        assert (a is None) == (b is None)
        if a is not None:
            assert sorted(m.from_ for m in a.round_change_messages) == b
        return a

    def valid_pc(self, cert, rlimit, height):
        a = self.o.valid_pc(cert, rlimit, height)
        b = self.c.valid_pc(ip.encode_pc(cert) if cert is not None else None, rlimit, height)
        assert a == b, (a, b)
        return a

    def validate_proposal(self, msg, h, r):
        v = ip.View(h, r)
        a = self.o.validate_proposal0(msg, v) if r == 0 else self.o.validate_proposal(msg, v)
        b = self.c.validate_proposal(enc(msg), h, r)
        assert a == b, (a, b)
        return a


def test_valid_pc_table_both_implementations():
    """core/ibft_test.go:1510-2015 TestIBFT_ValidPC (the wire-representable sub-cases) through oracle and C++."""
    n = 4
    is_sender = dict(is_proposer=lambda p, h, r: p == SENDER)
    mk = lambda **kw: Dual(n, **kw)  # noqa: E731  (validators node 0..3: the PC's three preparers already carry quorum 3)
    assert mk().valid_pc(None, 0, 0) is True
    assert mk().valid_pc(ip.PreparedCertificate(None, None), 0, 0) is False
    assert mk().valid_pc(ip.PreparedCertificate(ip.IbftMessage(view=ip.View()), None), 0, 0) is False
    c = ip.PreparedCertificate(ip.IbftMessage(view=ip.View(), type=ip.PREPARE, payload=ip.PrepareMessage()), gen_unique(n - 1, ip.PREPARE))
    assert mk().valid_pc(c, 1, 0) is False                                            # invalid proposal message type
    c, _ = pc_all_same()
    c.prepare_messages[0].type = ip.ROUND_CHANGE
    assert mk(**is_sender).valid_pc(c, 1, 0) is False                                 # invalid prepare message type
    c = ip.PreparedCertificate(gen_with_sender(1, ip.PREPREPARE, b"node x")[0], gen_with_sender(n - 1, ip.PREPARE, b"node x"))
    assert mk().valid_pc(c, 1, 0) is False                                            # non unique senders
    proposal = gen_with_sender(1, ip.PREPREPARE, SENDER)[0]
    c = ip.PreparedCertificate(proposal, gen_unique(n - 1, ip.PREPARE))
    append_hash([c.proposal_message], b"proposal hash 1")
    append_hash(c.prepare_messages, b"proposal hash 2")
    assert mk(**is_sender).valid_pc(c, 1, 0) is False                                 # differing proposal hashes
    c, _ = pc_all_same(rlimit=1, rnd=2)
    assert mk(**is_sender).valid_pc(c, 1, 0) is False                                 # rounds not lower than rLimit
    c, _ = pc_all_same()
    c.proposal_message.view.height = 10
    assert mk(**is_sender).valid_pc(c, 1, 0) is False                                 # heights are not the same
    c, _ = pc_all_same(rlimit=2)
    c.prepare_messages[1].view.round = 0
    assert mk(**is_sender).valid_pc(c, 2, 0) is False                                 # rounds are not the same
    c, _ = pc_all_same()
    assert mk(is_proposer=lambda p, h, r: p != SENDER).valid_pc(c, 1, 0) is False     # proposal not from proposer
    assert mk(is_valid_validator=lambda m: m.from_ != b"node 1", **is_sender).valid_pc(c, 1, 0) is False
    assert mk(is_valid_validator=lambda m: m.from_ != SENDER, **is_sender).valid_pc(c, 1, 0) is False
    assert mk(is_proposer=lambda p, h, r: True).valid_pc(c, 1, 0) is False            # prepare from proposer
    assert mk(is_valid_validator=lambda m: True, **is_sender).valid_pc(c, 1, 0) is True  # completely valid PC
    # uninitialised validator manager (the state the reference's negative sub-cases actually run in)
    assert Dual(n, init=False, **is_sender).valid_pc(c, 1, 0) is False


def test_is_acceptable_and_add_message_signal_matrix():
    """core/ibft_test.go:1103-1216 and :3120-3246 (TestIBFT_AddMessage: add/signal matrix)."""
    cases = [(None, (0, 0), True, False), (None, (0, 0), False, False), ((100, 0), (0, 0), False, True), ((100, 0), (0, 1), False, True),
             ((0, 100), (0, 0), False, True), ((0, 0), (0, 1), False, False), ((0, 0), (1, 0), False, False), ((0, 1), (1, 0), False, False)]
    for mview, cur, invalid_sender, want in cases:
        d = Dual(is_valid_validator=lambda m, inv=invalid_sender: not inv)
        d.set_state(*cur)
        msg = ip.IbftMessage(view=ip.View(*mview) if mview else None, type=ip.PREPARE, payload=ip.PrepareMessage(b"h"))
        assert d.o.is_acceptable_message(msg) == want == d.c.is_acceptable(enc(msg))
    # AddMessage: quorum of COMMITs at the state's height signals once the quorum-th distinct sender arrives
    d = Dual(4)
    d.set_state(1, 0)
    for i in range(4):
        d.add_message(ip.IbftMessage(ip.View(1, 0), b"node %d" % i, b"", ip.COMMIT, ip.CommitMessage(HASH, b"s")))
        assert d.c.signal_count() == len(d.o.messages.signals) == max(0, i - 1)
    d.add_message(ip.IbftMessage(ip.View(2, 0), b"node 0", b"", ip.COMMIT, ip.CommitMessage(HASH, b"s")))   # future height: stored, no signal
    assert d.c.signal_count() == len(d.o.messages.signals) == 2 and d.c.num_messages(2, 0, ip.COMMIT) == 1
    d.add_message(ip.IbftMessage(ip.View(0, 0), b"node 0", b"", ip.COMMIT, ip.CommitMessage(HASH, b"s")))   # past height: dropped
    assert d.c.num_messages(0, 0, ip.COMMIT) == 0


def test_prepare_commit_flow_with_pruning():
    """core/ibft_test.go:870-1099 shapes + messages_test.go:183-268 pruning, both implementations."""
    d = Dual(4, is_valid_proposal_hash=lambda p, h: h == HASH,
             is_valid_committed_seal=lambda h, s: s is not None and s.signature == b"good")
    proposal = ip.IbftMessage(ip.View(0, 0), b"node 0", b"", ip.PREPREPARE, ip.PrePrepareMessage(ip.Proposal(b"block", 0), HASH))
    d.set_state(0, 0, L.PREPARE_STATE, proposal)
    d.store_add(ip.IbftMessage(ip.View(0, 0), b"node 1", b"", ip.PREPARE, ip.PrepareMessage(HASH)))
    d.store_add(ip.IbftMessage(ip.View(0, 0), b"node 3", b"", ip.PREPARE, ip.PrepareMessage(b"bad")))
    assert d.handle_prepare(0, 0) is False                      # proposer + 1 < quorum 3; the bad-hash PREPARE is pruned
    assert d.c.num_messages(0, 0, ip.PREPARE) == 1
    d.store_add(ip.IbftMessage(ip.View(0, 0), b"node 2", b"", ip.PREPARE, ip.PrepareMessage(HASH)))
    assert d.handle_prepare(0, 0) is True
    for k, seal in ((0, b"good"), (1, b"good"), (2, b"bad"), (3, b"good")):
        d.store_add(ip.IbftMessage(ip.View(0, 0), b"node %d" % k, b"", ip.COMMIT, ip.CommitMessage(HASH, seal)))
    d.store_add(ip.IbftMessage(ip.View(0, 0), b"node 9", b"", ip.COMMIT, ip.PrepareMessage(HASH)))   # payload/type mismatch -> nil hash & seal
    assert d.handle_commit(0, 0) is True
    assert d.c.seal_count() == 3 and d.c.num_messages(0, 0, ip.COMMIT) == 3
    # proposer among the preparers
    d2 = Dual(4)
    d2.set_state(0, 0, L.PREPARE_STATE, proposal)
    for k in range(4):
        d2.store_add(ip.IbftMessage(ip.View(0, 0), b"node %d" % k, b"", ip.PREPARE, ip.PrepareMessage(HASH)))
    assert d2.handle_prepare(0, 0) is False


def filled_rc_messages(quorum, proposal, phash, rnd=2, pc_round=1):
    """generateFilledRCMessages, core/ibft_test.go:158-214."""
    rcs = gen_unique(quorum, ip.ROUND_CHANGE)
    prepares = gen_messages(quorum - 1, ip.PREPARE)
    for i, m in enumerate(prepares):
        m.payload = ip.PrepareMessage(phash)
        m.view = ip.View(0, pc_round)
        m.from_ = b"node %d" % (i + 1)
    pc = ip.PreparedCertificate(ip.IbftMessage(ip.View(0, pc_round), b"unique node", b"", ip.PREPREPARE,
                                               ip.PrePrepareMessage(proposal, phash, None)), prepares)
    for m in rcs:
        m.view = ip.View(0, rnd)
        m.payload = ip.RoundChangeMessage(proposal, pc)
    return rcs


def test_validate_proposal_table():
    """core/ibft_test.go:2017-2797 TestIBFT_ValidateProposal shapes, incl. the max-round hash rule (:2663-2796)."""
    q = 4
    raw = b"raw block"
    hash_of = lambda p: b"%s_%d" % (p.raw_proposal, p.round)  # noqa: E731  (ibft_test.go:2679-2701 uses fmt.Sprintf("%s_%d"))
    common = dict(is_proposer=lambda p, h, r: p == b"proposer", is_valid_proposal_hash=lambda p, h: p is not None and h == hash_of(p),
                  node_id=b"me")

    def pp(rnd, rcc, proposal_round=None, phash=None):
        prop = ip.Proposal(raw, rnd if proposal_round is None else proposal_round)
        return ip.IbftMessage(ip.View(0, rnd), b"proposer", b"", ip.PREPREPARE, ip.PrePrepareMessage(prop, phash or hash_of(prop), rcc))
    # round 0
    assert Dual(q, **common).validate_proposal(pp(0, None), 0, 0) is True
    assert Dual(q, **common).validate_proposal(pp(0, None, proposal_round=1), 0, 0) is False         # proposal round mismatch
    assert Dual(q, **common).validate_proposal(pp(0, None, phash=b"x"), 0, 0) is False                # hash mismatch
    assert Dual(q, **{**common, "is_valid_proposal": lambda r: False}).validate_proposal(pp(0, None), 0, 0) is False
    assert Dual(q, **{**common, "node_id": b"proposer"}).validate_proposal(pp(0, None), 0, 0) is False  # we are the proposer
    # round 2 with certificates
    def rcc_of(pc_round=1, pc_hash_round=None, n_rc=q):
        inner = ip.Proposal(raw, pc_round if pc_hash_round is None else pc_hash_round)
        return ip.RoundChangeCertificate(filled_rc_messages(n_rc, ip.Proposal(raw, pc_round), hash_of(inner), rnd=2, pc_round=pc_round))
    pc_ok = dict(common, is_proposer=lambda p, h, r: p in (b"proposer", b"unique node"))
    assert Dual(q, **pc_ok).validate_proposal(pp(2, None), 0, 2) is False                               # no certificate
    assert Dual(q, **pc_ok).validate_proposal(pp(2, ip.RoundChangeCertificate([])), 0, 2) is False       # empty RCC: no unique senders
    assert Dual(q, **pc_ok).validate_proposal(pp(2, rcc_of(n_rc=2)), 0, 2) is False                      # no RCC quorum
    assert Dual(q, **pc_ok).validate_proposal(pp(2, rcc_of()), 0, 2) is True                             # valid: hash(EB, maxR=1) matches
    assert Dual(q, **pc_ok).validate_proposal(pp(2, rcc_of(pc_hash_round=0)), 0, 2) is False             # max-round hash rule violated
    bad = rcc_of()
    bad.round_change_messages[1].view = ip.View(0, 1)
    assert Dual(q, **pc_ok).validate_proposal(pp(2, bad), 0, 2) is False                                 # RC round mismatch
    bad = rcc_of()
    bad.round_change_messages[2].type = ip.PREPARE
    assert Dual(q, **pc_ok).validate_proposal(pp(2, bad), 0, 2) is False                                 # wrong type inside RCC
    bad = rcc_of()
    bad.round_change_messages[3].from_ = bad.round_change_messages[0].from_
    assert Dual(q, **pc_ok).validate_proposal(pp(2, bad), 0, 2) is False                                 # duplicate sender
    assert Dual(q, **{**pc_ok, "is_valid_validator": lambda m: not (m.type == ip.ROUND_CHANGE and m.from_ == b"node 2")}
                ).validate_proposal(pp(2, rcc_of()), 0, 2) is False                                      # invalid RC sender
    # invalid PCs are skipped, not fatal: with every PC invalid the proposal is accepted on the RCC alone (ibft.go:763-765)
    assert Dual(q, **{**pc_ok, "is_valid_validator": lambda m: m.type != ip.PREPARE}).validate_proposal(pp(2, rcc_of(pc_hash_round=0)), 0, 2) is True


def test_round_change_extended_rcc():
    """messages_test.go:273-329 + core/ibft.go:470-512 through both implementations."""
    raw = b"blk"
    hash_of = lambda p: b"%s_%d" % (p.raw_proposal, p.round)  # noqa: E731
    d = Dual(4, is_proposer=lambda p, h, r: p == b"unique node", is_valid_proposal_hash=lambda p, h: p is not None and h == hash_of(p),
             extra_validators=(b"unique node",))
    d.set_state(0, 1)
    for m in filled_rc_messages(4, ip.Proposal(raw, 1), hash_of(ip.Proposal(raw, 1)), rnd=2, pc_round=1):
        d.store_add(m)
    for m in gen_unique(2, ip.ROUND_CHANGE):                      # round 3: below quorum
        m.view = ip.View(0, 3)
        d.store_add(m)
    bad = filled_rc_messages(4, ip.Proposal(raw, 1), b"wrong", rnd=4, pc_round=1)   # round 4: certificate does not match the proposal
    for m in bad:
        d.store_add(m)
    got = d.handle_round_change(0, 1)
    assert got is not None and {m.view.round for m in got.round_change_messages} == {2}
    assert d.c.num_messages(0, 4, ip.ROUND_CHANGE) == 4          # GetExtendedRCC does not prune
    d.set_state(0, 2, proposal=ip.IbftMessage(ip.View(0, 2), b"p", b"", ip.PREPREPARE, ip.PrePrepareMessage(ip.Proposal(raw, 2), b"h")))
    assert d.handle_round_change(0, 2) is None                   # same round and a proposal already accepted (ibft.go:492-494)


def test_randomised_differential_commit_rounds():
    """Random COMMIT/PREPARE traffic with pseudo-random verdicts: decisions, pruned stores and seal counts must agree."""
    rnd = random.Random(42)
    for trial in range(25):
        n = rnd.randint(4, 12)
        bad_seal = {b"node %d" % i for i in range(n) if rnd.random() < 0.2}
        bad_hash = {b"node %d" % i for i in range(n) if rnd.random() < 0.15}
        powers = {b"node %d" % i: rnd.randint(1, 5) for i in range(n)}
        d = Dual(n, powers=powers, is_valid_proposal_hash=lambda p, h: h == HASH,
                 is_valid_committed_seal=lambda h, s, b=bad_seal: s is not None and s.signer not in b)
        proposal = ip.IbftMessage(ip.View(7, 0), b"node 0", b"", ip.PREPREPARE, ip.PrePrepareMessage(ip.Proposal(b"blk", 0), HASH))
        d.set_state(7, 0, L.PREPARE_STATE, proposal)
        order = list(range(n))
        rnd.shuffle(order)
        for i in order[: rnd.randint(0, n)]:
            a = b"node %d" % i
            d.store_add(ip.IbftMessage(ip.View(7, 0), a, b"", ip.PREPARE, ip.PrepareMessage(b"bad" if a in bad_hash else HASH)))
        d.handle_prepare(7, 0)
        for i in order[: rnd.randint(0, n)]:
            a = b"node %d" % i
            d.store_add(ip.IbftMessage(ip.View(7, 0), a, b"", ip.COMMIT, ip.CommitMessage(b"bad" if a in bad_hash else HASH, b"seal")))
            if rnd.random() < 0.3:
                d.handle_commit(7, 0)
        d.handle_commit(7, 0)
        d.c.prune_by_height(8)
        assert d.c.num_messages(7, 0, ip.COMMIT) == 0


def test_incremental_quorum_accumulators_match_reference_semantics():
    """AddMessage with O(log N) incremental voting-power accumulators must signal exactly when the
    reference's recompute-everything path (core/ibft.go:1113-1120) does -- under replacement, pruning, table changes, the
    PREPARE proposer rule, unknown senders."""
    rnd = random.Random(77)
    for trial in range(30):
        n = rnd.randint(4, 10)
        powers = {b"node %d" % i: rnd.randint(1, 4) for i in range(n)}
        d = Dual(n, powers=powers, is_valid_proposal_hash=lambda p, h: h == HASH)
        d.c.set_incremental_quorum(True)
        proposal = ip.IbftMessage(ip.View(3, 0), b"node 0", b"", ip.PREPREPARE, ip.PrePrepareMessage(ip.Proposal(b"b", 0), HASH))
        d.set_state(3, 0, L.PREPARE_STATE if rnd.random() < 0.7 else L.NEW_ROUND, proposal if rnd.random() < 0.8 else None)
        for step in range(40):
            who = b"node %d" % rnd.randint(0, n + 1)                  # includes two non-validators
            t = rnd.choice([ip.PREPARE, ip.COMMIT, ip.ROUND_CHANGE, ip.PREPREPARE])
            view = ip.View(3, rnd.choice([0, 0, 0, 1]))
            payload = {ip.PREPARE: ip.PrepareMessage(HASH if rnd.random() < 0.8 else b"x"), ip.COMMIT: ip.CommitMessage(HASH, b"s"),
                       ip.ROUND_CHANGE: ip.RoundChangeMessage(), ip.PREPREPARE: ip.PrePrepareMessage(ip.Proposal(b"b", 0), HASH)}[t]
            d.add_message(ip.IbftMessage(view, who, bytes([step]), t, payload))
            assert d.c.signal_count() == len(d.o.messages.signals), (trial, step)
            if rnd.random() < 0.15:
                d.handle_prepare(3, 0)                                   # prunes bad-hash PREPAREs: accumulators must follow
                d.set_state(3, 0, L.PREPARE_STATE, proposal)
            if rnd.random() < 0.05:                                      # validator table changes (new epoch)
                powers = {k: rnd.randint(1, 4) for k in powers}
                d.vm.set_current_voting_power(powers)
                names = sorted(powers)
                assert d.c.set_validators(3, names, [powers[k] for k in names]) == 0


def test_incremental_quorum_is_faster_on_a_large_round():
    import time
    n = 3000
    addrs = [b"v%05d" % i for i in range(n)]
    wires = [enc(ip.IbftMessage(ip.View(1, 0), a, b"", ip.COMMIT, ip.CommitMessage(HASH, b"s"))) for a in addrs]
    times = {}
    for inc in (False, True):
        c = host.HostContext("callback")
        assert c.set_validators(1, addrs, None) == 0
        c.set_state(1, 0)
        c.set_incremental_quorum(inc)
        t0 = time.perf_counter()
        c.add_messages(wires)
        times[inc] = time.perf_counter() - t0
        assert c.signal_count() == n - (2 * n // 3 + 1) + 1          # one signal per arrival from the quorum-th sender on
    assert times[True] * 5 < times[False], times
