"""The reference's own decision tables for the host logic around the hot path, replayed against the restatement
oracle/ibft_logic.py.  Each test cites the reference test it restates (file:line under /root/reference)."""
import pytest

from oracle import ibft_logic as L
from oracle import ibft_proto as ip

# core/validator_manager_test.go:18-187 -- (voting powers, signers, hasQuorum)
QUORUM_CASES = [
    ({"A": 1, "B": 1, "C": 1, "D": 1}, "ABCD", True),
    ({"A": 1, "B": 1, "C": 1, "D": 1}, "AB", False),
    ({"A": 1, "B": 1, "C": 1, "D": 1, "E": 1, "F": 1}, "ABCDE", True),
    ({"A": 1, "B": 1, "C": 1, "D": 1, "E": 1, "F": 1}, "ABCD", False),
    ({"A": 2, "B": 2, "C": 2, "D": 3}, "ACD", True),
    ({"A": 2, "B": 2, "C": 2, "D": 3}, "AD", False),
    ({"A": 2, "B": 2, "C": 3, "D": 3}, "ABD", True),
    ({"A": 2, "B": 2, "C": 3, "D": 3}, "AD", False),
    ({"A": 2, "B": 7, "C": 7, "D": 5}, "ABC", True),
    ({"A": 2, "B": 7, "C": 7, "D": 5}, "CD", False),
]


@pytest.mark.parametrize("powers,signers,want", QUORUM_CASES)
def test_calculate_quorum_table(powers, signers, want):
    vm = L.ValidatorManager(lambda h: {})
    vm.set_current_voting_power({k.encode(): v for k, v in powers.items()})
    assert vm.has_quorum({s.encode() for s in signers}) == want


def test_quorum_uninitialised_and_zero_power():
    vm = L.ValidatorManager(lambda h: {})
    assert vm.has_quorum({b"A"}) is False                       # validator_manager.go:82-84
    with pytest.raises(L.VotingPowerError):
        vm.set_current_voting_power({b"A": 0})                  # validator_manager.go:66-68
    assert [L.calculate_quorum(t) for t in (4, 6, 9, 10, 21)] == [3, 5, 7, 7, 15]


# ---- generators of core/ibft_test.go:55-134
def gen_messages(count, mtype):
    payload = {ip.PREPREPARE: ip.PrePrepareMessage, ip.PREPARE: ip.PrepareMessage, ip.COMMIT: ip.CommitMessage,
               ip.ROUND_CHANGE: ip.RoundChangeMessage}[mtype]
    return [ip.IbftMessage(ip.View(0, 0), b"", b"", mtype, payload()) for _ in range(count)]


def gen_with_sender(count, mtype, sender):
    ms = gen_messages(count, mtype)
    for m in ms:
        m.from_ = sender
    return ms


def gen_unique(count, mtype):
    ms = gen_messages(count, mtype)
    for i, m in enumerate(ms):
        m.from_ = b"node %d" % i
    return ms


def append_hash(ms, h):
    for m in ms:
        if m.type in (ip.PREPREPARE, ip.PREPARE):
            m.payload.proposal_hash = h


def set_round(ms, r):
    for m in ms:
        m.view.round = r


def vp_for_cnt(n):
    return lambda h: {b"node %d" % i: 1 for i in range(n)}


def make_ibft(backend=None, n=4, init=False):
    vm = L.ValidatorManager(vp_for_cnt(n))
    if init:
        vm.init(0)
    return L.IBFT(backend or L.Backend(), vm)


HASH = b"proposal hash"
SENDER = b"unique node"


def pc_all_same(quorum=4, rlimit=1, rnd=None):
    proposal = gen_with_sender(1, ip.PREPREPARE, SENDER)[0]
    cert = ip.PreparedCertificate(proposal, gen_unique(quorum - 1, ip.PREPARE))
    allm = [cert.proposal_message] + cert.prepare_messages
    append_hash(allm, HASH)
    set_round(allm, rlimit - 1 if rnd is None else rnd)
    return cert, allm


@pytest.mark.parametrize("init", [False, True])
def test_valid_pc_table(init):
    """core/ibft_test.go:1510-2015 TestIBFT_ValidPC, 15 sub-cases.  The reference initialises the validator manager only in
    the last sub-case; with init=True every earlier case is isolated to the check its name describes (the proposer is made a
    validator so that the quorum test passes)."""
    n = 4
    not_sender = L.Backend(is_proposer=lambda p, h, r: p != SENDER)
    is_sender = lambda **kw: L.Backend(is_proposer=lambda p, h, r: p == SENDER, **kw)  # noqa: E731

    def mk(backend=None):
        i = make_ibft(backend, n, init)
        if init:
            i.vm.set_current_voting_power({**vp_for_cnt(n)(0), SENDER: 1, b"node x": 1})
        return i

    assert mk().valid_pc(None, 0, 0) is True                                             # "no certificate"
    assert mk().valid_pc(ip.PreparedCertificate(None, []), 0, 0) is False                # "proposal and prepare messages mismatch"
    assert mk().valid_pc(ip.PreparedCertificate(ip.IbftMessage(), None), 0, 0) is False
    # "no Quorum PP + P messages": quorum-2 prepares with empty senders
    assert mk().valid_pc(ip.PreparedCertificate(ip.IbftMessage(), gen_messages(n - 2, ip.PREPARE)), 0, 0) is False
    # "invalid proposal message type"
    assert mk().valid_pc(ip.PreparedCertificate(ip.IbftMessage(type=ip.PREPARE), gen_messages(n - 1, ip.PREPARE)), 0, 0) is False
    # "invalid prepare message type"
    c = ip.PreparedCertificate(ip.IbftMessage(type=ip.PREPREPARE), gen_messages(n - 1, ip.PREPARE))
    c.prepare_messages[0].type = ip.ROUND_CHANGE
    assert mk().valid_pc(c, 0, 0) is False
    # "non unique senders"
    c = ip.PreparedCertificate(ip.IbftMessage(view=ip.View(), type=ip.PREPREPARE, from_=b"node x", payload=ip.PrePrepareMessage()),
                               gen_with_sender(n - 1, ip.PREPARE, b"node x"))
    assert mk().valid_pc(c, 0, 0) is False
    # "differing proposal hashes"
    proposal = gen_with_sender(1, ip.PREPREPARE, SENDER)[0]
    c = ip.PreparedCertificate(proposal, gen_unique(n - 1, ip.PREPARE))
    append_hash([c.proposal_message], b"proposal hash 1")
    append_hash(c.prepare_messages, b"proposal hash 2")
    assert mk(is_sender()).valid_pc(c, 1, 0) is False
    # "rounds not lower than rLimit"
    c, _ = pc_all_same(rlimit=1, rnd=2)
    assert mk(is_sender()).valid_pc(c, 1, 0) is False
    # "heights are not the same"
    c, _ = pc_all_same(rlimit=1)
    c.proposal_message.view.height = 10
    assert mk(is_sender()).valid_pc(c, 1, 0) is False
    # "rounds are not the same"
    c, _ = pc_all_same(rlimit=2)
    c.prepare_messages[1].view.round = 0
    assert mk(is_sender()).valid_pc(c, 2, 0) is False
    # "proposal not from proposer"
    c, _ = pc_all_same()
    assert mk(not_sender).valid_pc(c, 1, 0) is False
    # "prepare is from an invalid sender"
    c, _ = pc_all_same()
    assert mk(is_sender(is_valid_validator=lambda m: m.from_ != b"node 1")).valid_pc(c, 1, 0) is False
    # "proposal is from an invalid sender"
    c, _ = pc_all_same()
    assert mk(is_sender(is_valid_validator=lambda m: m.from_ != SENDER)).valid_pc(c, 1, 0) is False
    # "prepare from proposer"
    c, _ = pc_all_same()
    assert mk(L.Backend(is_proposer=lambda p, h, r: True)).valid_pc(c, 1, 0) is False
    # "completely valid PC" (the reference calls validatorManager.Init(0) here: ibft_test.go:1993)
    c, _ = pc_all_same()
    i = make_ibft(is_sender(is_valid_validator=lambda m: True), n, init=True)
    assert i.valid_pc(c, 1, 0) is True


def test_is_acceptable_message_table():
    """core/ibft_test.go:1103-1216 TestIBFT_IsAcceptableMessage."""
    base_view = ip.View(0, 0)
    cases = [  # (name, msgView, currentView, invalidSender, acceptable)
        ("invalid sender", None, base_view, True, False),
        ("malformed message", None, base_view, False, False),
        ("higher height, same round number", ip.View(100, 0), base_view, False, True),
        ("higher height, lower round number", ip.View(100, 0), ip.View(0, 1), False, True),
        ("same heights, higher round number", ip.View(0, 100), base_view, False, True),
        ("same heights, lower round number", ip.View(0, 0), ip.View(0, 1), False, False),
        ("lower height, same round number", ip.View(0, 0), ip.View(1, 0), False, False),
        ("lower height, higher round number", ip.View(0, 1), ip.View(1, 0), False, False),
    ]
    for name, mview, cur, invalid_sender, want in cases:
        i = make_ibft(L.Backend(is_valid_validator=lambda m, inv=invalid_sender: not inv))
        i.state.view = ip.View(cur.height, cur.round)
        assert i.is_acceptable_message(ip.IbftMessage(view=mview, type=ip.PREPARE)) == want, name


def test_store_prunes_invalid_and_dedups():
    """messages/messages_test.go:100-128 (dedup: last write wins per sender) and :183-268 (GetValidMessages prunes)."""
    ms = L.Messages()
    v = ip.View(1, 0)
    for i in range(5):
        ms.add_message(ip.IbftMessage(v, b"node %d" % i, b"", ip.PREPARE, ip.PrepareMessage(b"h")))
    ms.add_message(ip.IbftMessage(v, b"node 0", b"sig2", ip.PREPARE, ip.PrepareMessage(b"h2")))
    assert ms.num_messages(v, ip.PREPARE) == 5
    got = ms.get_valid_messages(v, ip.PREPARE, lambda m: m.from_ != b"node 3")
    assert sorted(m.from_ for m in got) == [b"node 0", b"node 1", b"node 2", b"node 4"]
    assert ms.num_messages(v, ip.PREPARE) == 4                       # invalid message pruned (messages.go:193-196)
    assert [m for m in got if m.from_ == b"node 0"][0].signature == b"sig2"
    ms.prune_by_height(2)
    assert ms.num_messages(v, ip.PREPARE) == 0


def test_extended_rcc_picks_highest_valid_round_and_does_not_prune():
    """messages/messages_test.go:273-329."""
    ms = L.Messages()
    for rnd, count in ((1, 4), (2, 4), (3, 2)):
        for i in range(count):
            ms.add_message(ip.IbftMessage(ip.View(0, rnd), b"node %d" % i, b"", ip.ROUND_CHANGE, ip.RoundChangeMessage()))
    ext = ms.get_extended_rcc(0, lambda m: True, lambda r, msgs: len(msgs) >= 3)
    assert ext is not None and {m.view.round for m in ext} == {2}
    assert ms.get_extended_rcc(0, lambda m: False, lambda r, msgs: len(msgs) >= 3) is None
    assert ms.num_messages(ip.View(0, 2), ip.ROUND_CHANGE) == 4      # GetExtendedRCC does not prune (messages.go:202-245)
    most = ms.get_most_round_change_messages(2, 0)
    assert len(most) == 4 and {m.view.round for m in most} == {2}
    assert ms.get_most_round_change_messages(4, 0) is None


def test_are_valid_pc_messages_and_unique_senders():
    """messages/helpers_test.go AreValidPCMessages / HasUniqueSenders tables (shapes)."""
    assert L.has_unique_senders([]) is False
    assert L.has_unique_senders(gen_unique(3, ip.PREPARE)) is True
    assert L.has_unique_senders(gen_with_sender(2, ip.PREPARE, b"x")) is False
    ms = gen_unique(3, ip.PREPARE)
    append_hash(ms, b"h")
    assert L.are_valid_pc_messages(ms, 0, 1) is True
    assert L.are_valid_pc_messages(ms, 0, 0) is False                # round >= roundLimit
    assert L.are_valid_pc_messages(ms, 1, 1) is False                # height
    assert L.are_valid_pc_messages([], 0, 1) is False
    ms[1].payload.proposal_hash = b"other"
    assert L.are_valid_pc_messages(ms, 0, 1) is False
    assert L.are_valid_pc_messages(gen_unique(2, ip.COMMIT), 0, 1) is False   # COMMIT / ROUND_CHANGE carry no PC hash


def test_extractors_return_none_on_mismatch():
    """messages/helpers.go:38-146: nil on type/payload mismatch (the verifier must then answer false, not crash)."""
    m = ip.IbftMessage(ip.View(), b"a", b"", ip.PREPARE, ip.CommitMessage(b"h", b"s"))
    assert L.extract_prepare_hash(m) is None and L.extract_commit_hash(m) is None
    assert L.extract_committed_seal(m) == L.CommittedSeal(b"a", b"s")   # only the payload type is checked (helpers.go:38-48)
    assert L.extract_committed_seal(ip.IbftMessage(type=ip.COMMIT)) is None
    assert L.extract_proposal(ip.IbftMessage(type=ip.PREPREPARE)) is None
    assert L.extract_latest_pc(ip.IbftMessage(type=ip.ROUND_CHANGE, payload=ip.RoundChangeMessage())) is None
    seals, err = L.extract_committed_seals([ip.IbftMessage(type=ip.PREPARE)])
    assert seals is None and err


def test_handle_prepare_and_commit_transitions():
    """core/ibft_test.go:870-1099 TestRunPrepare / TestRunCommit shapes: quorum of valid messages moves the state."""
    n = 4
    vm = L.ValidatorManager(vp_for_cnt(n))
    vm.init(0)
    i = L.IBFT(L.Backend(is_valid_proposal_hash=lambda p, h: h == HASH), vm)
    v = ip.View(0, 0)
    i.state.proposal_message = ip.IbftMessage(v, b"node 0", b"", ip.PREPREPARE, ip.PrePrepareMessage(ip.Proposal(b"block", 0), HASH))
    i.state.name = L.PREPARE_STATE
    for k in (1, 2):
        i.messages.add_message(ip.IbftMessage(v, b"node %d" % k, b"", ip.PREPARE, ip.PrepareMessage(HASH)))
    i.messages.add_message(ip.IbftMessage(v, b"node 3", b"", ip.PREPARE, ip.PrepareMessage(b"bad")))
    assert i.handle_prepare(v) is True and i.state.name == L.COMMIT_STATE       # proposer + 2 preparers = 3 >= quorum 3
    assert len(i.state.latest_pc.prepare_messages) == 2
    for k in range(2):
        i.messages.add_message(ip.IbftMessage(v, b"node %d" % k, b"", ip.COMMIT, ip.CommitMessage(HASH, b"seal")))
    assert i.handle_commit(v) is False
    i.messages.add_message(ip.IbftMessage(v, b"node 2", b"", ip.COMMIT, ip.CommitMessage(HASH, b"seal")))
    assert i.handle_commit(v) is True and i.state.name == L.FIN_STATE and len(i.state.seals) == 3
    # proposer among the preparers => no prepare quorum (validator_manager.go:116-121)
    i2 = L.IBFT(L.Backend(), vm)
    i2.state.proposal_message = i.state.proposal_message
    for k in (0, 1, 2, 3):
        i2.messages.add_message(ip.IbftMessage(v, b"node %d" % k, b"", ip.PREPARE, ip.PrepareMessage(HASH)))
    assert i2.handle_prepare(v) is False


def test_are_valid_pc_messages_nil_hash_quirk():
    """helpers.go:190-198: `if hash == nil { hash = extractedHash }` -- an empty (nil) hash never becomes the reference, so
    [empty, "h"] passes while ["h", empty] fails."""
    a, b = gen_unique(2, ip.PREPARE)
    b.payload.proposal_hash = b"h"
    assert L.are_valid_pc_messages([a, b], 0, 1) is True
    assert L.are_valid_pc_messages([b, a], 0, 1) is False
