"""Restatement of the reference's host logic AROUND the hot path (plain Python, test infrastructure only).

Follows, function by function:
  core/validator_manager.go:23-155   ValidatorManager (HasQuorum, HasPrepareQuorum, calculateQuorum, ...)
  messages/messages.go:54-323        Messages store (AddMessage, GetValidMessages [prunes], GetExtendedRCC [does not prune],
                                     GetMostRoundChangeMessages, PruneByHeight)
  messages/helpers.go:16-227         payload extractors (nil on mismatch), HasUniqueSenders, AreValidPCMessages
  core/ibft.go:470-551, 629-813, 855-967, 1101-1149, 1162-1231, 1273-1284
                                     handleRoundChangeMessage, proposalMatchesCertificate, validateProposal*, handlePrePrepare,
                                     handlePrepare, handleCommit, AddMessage / isAcceptableMessage, validPC, hasQuorumByMsgType
The reference's own test tables for these functions (core/validator_manager_test.go, core/ibft_test.go TestIBFT_ValidPC /
TestIBFT_ValidateProposal / TestIBFT_IsAcceptableMessage, messages/messages_test.go, messages/helpers_test.go) are replayed
against this module in tests/test_oracle_logic.py; the C++ host mirror is then compared with it on the same inputs.

Message model: oracle/ibft_proto.py dataclasses (None == Go nil).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

from . import ibft_proto as ip

# stateType, core/state.go:10-18
NEW_ROUND, PREPARE_STATE, COMMIT_STATE, FIN_STATE = 0, 1, 2, 3


# ----------------------------------------------------------------------------- messages/helpers.go
@dataclass
class CommittedSeal:  # helpers.go:16-19
    signer: bytes
    signature: bytes


def extract_committed_seal(m: ip.IbftMessage) -> Optional[CommittedSeal]:  # helpers.go:38-48
    if not isinstance(m.payload, ip.CommitMessage):
        return None
    return CommittedSeal(m.from_, m.payload.committed_seal)


def extract_committed_seals(msgs: List[ip.IbftMessage]):  # helpers.go:22-35
    out = []
    for m in msgs:
        if m.type != ip.COMMIT:
            return None, "wrong type message is included in COMMIT messages"
        out.append(extract_committed_seal(m))
    return out, None


def extract_commit_hash(m: ip.IbftMessage) -> Optional[bytes]:  # helpers.go:51-62
    if m.type != ip.COMMIT or not isinstance(m.payload, ip.CommitMessage):
        return None
    return m.payload.proposal_hash


def extract_proposal(m: ip.IbftMessage) -> Optional[ip.Proposal]:  # helpers.go:65-76
    if m.type != ip.PREPREPARE or not isinstance(m.payload, ip.PrePrepareMessage):
        return None
    return m.payload.proposal


def extract_proposal_hash(m: ip.IbftMessage) -> Optional[bytes]:  # helpers.go:79-90
    if m.type != ip.PREPREPARE or not isinstance(m.payload, ip.PrePrepareMessage):
        return None
    return m.payload.proposal_hash


def extract_round_change_certificate(m: ip.IbftMessage) -> Optional[ip.RoundChangeCertificate]:  # helpers.go:93-104
    if m.type != ip.PREPREPARE or not isinstance(m.payload, ip.PrePrepareMessage):
        return None
    return m.payload.certificate


def extract_prepare_hash(m: ip.IbftMessage) -> Optional[bytes]:  # helpers.go:107-118
    if m.type != ip.PREPARE or not isinstance(m.payload, ip.PrepareMessage):
        return None
    return m.payload.proposal_hash


def extract_latest_pc(m: ip.IbftMessage) -> Optional[ip.PreparedCertificate]:  # helpers.go:121-132
    if m.type != ip.ROUND_CHANGE or not isinstance(m.payload, ip.RoundChangeMessage):
        return None
    return m.payload.latest_prepared_certificate


def extract_last_prepared_proposal(m: ip.IbftMessage) -> Optional[ip.Proposal]:  # helpers.go:135-146
    if m.type != ip.ROUND_CHANGE or not isinstance(m.payload, ip.RoundChangeMessage):
        return None
    return m.payload.last_prepared_proposal


def has_unique_senders(msgs: List[ip.IbftMessage]) -> bool:  # helpers.go:149-166
    if len(msgs) < 1:
        return False
    seen = set()
    for m in msgs:
        if m.from_ in seen:
            return False
        seen.add(m.from_)
    return True


def _bytes_equal(a: Optional[bytes], b: Optional[bytes]) -> bool:
    """Go bytes.Equal: nil and empty compare equal."""
    return (a or b"") == (b or b"")


def _extract_pc_message_hash(m: ip.IbftMessage):  # helpers.go:216-227
    if m.type == ip.PREPREPARE:
        return extract_proposal_hash(m), True
    if m.type == ip.PREPARE:
        return extract_prepare_hash(m), True
    return None, False


def are_valid_pc_messages(msgs: List[ip.IbftMessage], height: int, round_limit: int) -> bool:  # helpers.go:169-213
    if len(msgs) < 1:
        return False
    rnd = msgs[0].view.round
    senders = set()
    h = None
    for m in msgs:
        if m.view.height != height:
            return False
        if m.view.round != rnd or m.view.round >= round_limit:
            return False
        extracted, ok = _extract_pc_message_hash(m)
        if not h:  # Go: `if hash == nil` -- a wire-decoded empty hash is a nil slice, so it never becomes the reference hash
            h = extracted
        if not ok or not _bytes_equal(h, extracted):
            return False
        if m.from_ in senders:
            return False
        senders.add(m.from_)
    return True


# ----------------------------------------------------------------------------- core/validator_manager.go
class VotingPowerError(Exception):  # errVotingPowerNotCorrect, validator_manager.go:12-14
    pass


def calculate_quorum(total: int) -> int:  # validator_manager.go:130-135
    return (2 * total) // 3 + 1


class ValidatorManager:
    def __init__(self, get_voting_powers: Callable[[int], Dict[bytes, int]]):
        self.backend_get = get_voting_powers
        self.quorum_size = 0
        self.voting_power: Optional[Dict[bytes, int]] = None
        self.errors: List[str] = []

    def init(self, height: int):  # :50-57
        self.set_current_voting_power(self.backend_get(height))

    def set_current_voting_power(self, vp: Dict[bytes, int]):  # :61-74
        total = sum(vp.values())
        if total <= 0:
            raise VotingPowerError("total voting power is zero or less")
        self.voting_power = vp
        self.quorum_size = calculate_quorum(total)

    def has_quorum(self, senders) -> bool:  # :77-96
        if self.voting_power is None:
            return False
        power = 0
        for a in set(senders):
            if a in self.voting_power:
                power += self.voting_power[a]
        return power >= self.quorum_size

    def has_prepare_quorum(self, state_name: int, proposal_message: Optional[ip.IbftMessage], msgs: List[ip.IbftMessage]) -> bool:  # :99-127
        if proposal_message is None:
            if state_name == PREPARE_STATE:
                self.errors.append("HasPrepareQuorum - proposalMessage is not set")
            return False
        proposer = proposal_message.from_
        senders = {proposer}
        for m in msgs:
            if _bytes_equal(m.from_, proposer):
                self.errors.append("HasPrepareQuorum - proposer is among signers but it is not expected to be")
                return False
            senders.add(m.from_)
        return self.has_quorum(senders)


def convert_message_to_address_set(msgs: List[ip.IbftMessage]):  # :147-155
    return {m.from_ for m in msgs}


# ----------------------------------------------------------------------------- messages/messages.go
class Messages:
    """height -> round -> sender -> message, one map per type (messages.go:289-296).  Last write wins per sender."""

    def __init__(self):
        self.maps = {t: {} for t in (ip.PREPREPARE, ip.PREPARE, ip.COMMIT, ip.ROUND_CHANGE)}
        self.signals: List[tuple] = []

    def add_message(self, m: ip.IbftMessage):  # :54-65
        self.maps[m.type].setdefault(m.view.height, {}).setdefault(m.view.round, {})[m.from_] = m

    def signal_event(self, msg_type: int, view: ip.View):  # :68-72
        self.signals.append((msg_type, view.height, view.round))

    def num_messages(self, view: ip.View, msg_type: int) -> int:  # :96-119
        return len(self.maps[msg_type].get(view.height, {}).get(view.round, {}))

    def prune_by_height(self, height: int):  # :123-148
        for mp in self.maps.values():
            for h in [h for h in mp if h < height]:
                del mp[h]

    def get_valid_messages(self, view: ip.View, msg_type: int, is_valid) -> List[ip.IbftMessage]:  # :169-199
        msgs = self.maps[msg_type].get(view.height, {}).get(view.round)
        if msgs is None:
            return []
        valid, invalid_keys = [], []
        for key, m in list(msgs.items()):
            if not is_valid(m):
                invalid_keys.append(key)
                continue
            valid.append(m)
        for k in invalid_keys:  # prune out invalid messages
            del msgs[k]
        return valid

    def get_extended_rcc(self, height: int, is_valid_message, is_valid_rcc) -> Optional[List[ip.IbftMessage]]:  # :202-245
        round_map = self.maps[ip.ROUND_CHANGE].get(height, {})
        highest, extended = 0, None
        for rnd, msgs in round_map.items():
            if rnd <= highest:
                continue
            valid = [m for m in msgs.values() if is_valid_message(m)]
            if not is_valid_rcc(rnd, valid):
                continue
            highest, extended = rnd, valid
        return extended

    def get_most_round_change_messages(self, min_round: int, height: int) -> Optional[List[ip.IbftMessage]]:  # :249-286
        round_map = self.maps[ip.ROUND_CHANGE].get(height, {})
        best_round, best_count = 0, 0
        for rnd, msgs in round_map.items():
            if rnd < min_round:
                continue
            if len(msgs) > best_count:
                best_round, best_count = rnd, len(msgs)
        if best_round == 0:
            return None
        return list(round_map[best_round].values())


# ----------------------------------------------------------------------------- core/ibft.go (validation half)
class Backend:
    """core.Verifier + ID (core/backend.go:37-56, :84).  Defaults mirror the reference's mockBackend
    (core/mock_test.go:105-151): verifier methods default true, IsProposer default false."""

    def __init__(self, **fns):
        self.fns = fns

    def is_valid_proposal(self, raw):
        f = self.fns.get("is_valid_proposal")
        return f(raw) if f else True

    def is_valid_validator(self, m):
        f = self.fns.get("is_valid_validator")
        return f(m) if f else True

    def is_proposer(self, ident, height, rnd):
        f = self.fns.get("is_proposer")
        return f(ident, height, rnd) if f else False

    def is_valid_proposal_hash(self, proposal, h):
        f = self.fns.get("is_valid_proposal_hash")
        return f(proposal, h) if f else True

    def is_valid_committed_seal(self, h, seal):
        f = self.fns.get("is_valid_committed_seal")
        return f(h, seal) if f else True

    def id(self):
        f = self.fns.get("id")
        return f() if f else b""


class State:  # the part of core/state.go the predicates read
    def __init__(self):
        self.view = ip.View(0, 0)
        self.proposal_message: Optional[ip.IbftMessage] = None
        self.name = NEW_ROUND
        self.latest_pc = None
        self.latest_prepared_proposal = None
        self.seals = []

    def get_proposal(self) -> Optional[ip.Proposal]:  # state.go:135-144
        if self.proposal_message is not None:
            return extract_proposal(self.proposal_message)
        return None

    def get_height(self):
        return self.view.height

    def get_round(self):
        return self.view.round


class IBFT:
    def __init__(self, backend: Backend, vm: ValidatorManager, messages: Optional[Messages] = None):
        self.backend, self.vm = backend, vm
        self.messages = messages or Messages()
        self.state = State()
        self.sent_commit = 0

    # core/ibft.go:1273-1284
    def has_quorum_by_msg_type(self, msgs, msg_type) -> bool:
        if msg_type == ip.PREPREPARE:
            return len(msgs) >= 1
        if msg_type == ip.PREPARE:
            return self.vm.has_prepare_quorum(self.state.name, self.state.proposal_message, msgs)
        if msg_type in (ip.ROUND_CHANGE, ip.COMMIT):
            return self.vm.has_quorum(convert_message_to_address_set(msgs))
        return False

    # core/ibft.go:1162-1231
    def valid_pc(self, cert: Optional[ip.PreparedCertificate], round_limit: int, height: int) -> bool:
        if cert is None:
            return True
        if cert.proposal_message is None or cert.prepare_messages is None:
            return False
        all_msgs = [cert.proposal_message] + list(cert.prepare_messages)
        if not self.vm.has_quorum(convert_message_to_address_set(all_msgs)):
            return False
        if cert.proposal_message.type != ip.PREPREPARE:
            return False
        for m in cert.prepare_messages:
            if m.type != ip.PREPARE:
                return False
        if not are_valid_pc_messages(all_msgs, height, round_limit):
            return False
        pm = cert.proposal_message
        if not self.backend.is_proposer(pm.from_, pm.view.height, pm.view.round):
            return False
        if not self.backend.is_valid_validator(pm):
            return False
        for m in cert.prepare_messages:
            if not self.backend.is_valid_validator(m):
                return False
            if self.backend.is_proposer(m.from_, m.view.height, m.view.round):
                return False
        return True

    # core/ibft.go:516-551
    def proposal_matches_certificate(self, proposal, cert) -> bool:
        if proposal is None and cert is None:
            return True
        if cert is None:
            return False
        hashes = [extract_proposal_hash(cert.proposal_message)]
        for m in cert.prepare_messages or []:
            hashes.append(extract_prepare_hash(m))
        for h in hashes:
            if not self.backend.is_valid_proposal_hash(proposal, h):
                return False
        return True

    # core/ibft.go:629-655
    def validate_proposal_common(self, msg, view) -> bool:
        proposal = extract_proposal(msg)
        proposal_hash = extract_proposal_hash(msg)
        if proposal.round != view.round:
            return False
        if not self.backend.is_proposer(msg.from_, view.height, view.round):
            return False
        if not self.backend.is_valid_proposal_hash(proposal, proposal_hash):
            return False
        return self.backend.is_valid_proposal(proposal.raw_proposal)

    # core/ibft.go:658-680
    def validate_proposal0(self, msg, view) -> bool:
        if msg.view.round != 0:
            return False
        if not self.validate_proposal_common(msg, view):
            return False
        if self.backend.is_proposer(self.backend.id(), view.height, view.round):
            return False
        return True

    # core/ibft.go:683-788
    def validate_proposal(self, msg, view) -> bool:
        height, rnd = view.height, view.round
        proposal = extract_proposal(msg)
        rcc = extract_round_change_certificate(msg)
        if not self.validate_proposal_common(msg, view):
            return False
        if rcc is None:
            return False
        if not has_unique_senders(rcc.round_change_messages):
            return False
        if not self.has_quorum_by_msg_type(rcc.round_change_messages, ip.ROUND_CHANGE):
            return False
        if self.backend.is_proposer(self.backend.id(), height, rnd):
            return False
        for rc in rcc.round_change_messages:
            if rc.type != ip.ROUND_CHANGE:
                return False
            if rc.view.height != height:
                return False
            if rc.view.round != rnd:
                return False
            if not self.backend.is_valid_validator(rc):
                return False
        tuples = []
        for rc in rcc.round_change_messages:
            cert = extract_latest_pc(rc)
            if cert is not None and self.valid_pc(cert, msg.view.round, height):
                tuples.append((cert.proposal_message.view.round, extract_proposal_hash(cert.proposal_message)))
        if not tuples:
            return True
        max_round, expected = 0, None
        for r, h in tuples:
            if r >= max_round:
                max_round, expected = r, h
        return self.backend.is_valid_proposal_hash(ip.Proposal(proposal.raw_proposal, max_round), expected)

    # core/ibft.go:792-813
    def handle_preprepare(self, view) -> Optional[ip.IbftMessage]:
        def is_valid(m):
            if view.round == 0:
                return self.validate_proposal0(m, view)
            return self.validate_proposal(m, view)
        msgs = self.messages.get_valid_messages(view, ip.PREPREPARE, is_valid)
        return msgs[0] if msgs else None

    # core/ibft.go:855-889
    def handle_prepare(self, view) -> bool:
        def is_valid(m):
            return self.backend.is_valid_proposal_hash(self.state.get_proposal(), extract_prepare_hash(m))
        prepares = self.messages.get_valid_messages(view, ip.PREPARE, is_valid)
        if not self.has_quorum_by_msg_type(prepares, ip.PREPARE):
            return False
        self.sent_commit += 1
        self.state.latest_pc = ip.PreparedCertificate(self.state.proposal_message, prepares)
        self.state.latest_prepared_proposal = self.state.get_proposal()
        self.state.name = COMMIT_STATE
        return True

    # core/ibft.go:931-967
    def handle_commit(self, view) -> bool:
        def is_valid(m):
            h = extract_commit_hash(m)
            seal = extract_committed_seal(m)
            if not self.backend.is_valid_proposal_hash(self.state.get_proposal(), h):
                return False
            return self.backend.is_valid_committed_seal(h, seal)
        commits = self.messages.get_valid_messages(view, ip.COMMIT, is_valid)
        if not self.has_quorum_by_msg_type(commits, ip.COMMIT):
            return False
        seals, err = extract_committed_seals(commits)
        if err:
            return False
        self.state.seals = seals
        self.state.name = FIN_STATE
        return True

    # core/ibft.go:470-512
    def handle_round_change_message(self, view) -> Optional[ip.RoundChangeCertificate]:
        height = view.height
        has_accepted = self.state.get_proposal() is not None

        def is_valid_msg(m):
            proposal = extract_last_prepared_proposal(m)
            cert = extract_latest_pc(m)
            if not self.valid_pc(cert, m.view.round, height):
                return False
            return self.proposal_matches_certificate(proposal, cert)

        def is_valid_rcc(rnd, msgs):
            if rnd == view.round and has_accepted:
                return False
            return self.has_quorum_by_msg_type(msgs, ip.ROUND_CHANGE)

        ext = self.messages.get_extended_rcc(height, is_valid_msg, is_valid_rcc)
        if ext is None:
            return None
        return ip.RoundChangeCertificate(ext)

    # core/ibft.go:1126-1149
    def is_acceptable_message(self, m) -> bool:
        if not self.backend.is_valid_validator(m):
            return False
        if m.view is None:
            return False
        if self.state.get_height() > m.view.height:
            return False
        if self.state.get_height() == m.view.height:
            return m.view.round >= self.state.get_round()
        return True

    # core/ibft.go:1101-1123
    def add_message(self, m):
        if m is None:
            return
        if self.is_acceptable_message(m):
            self.messages.add_message(m)
            if m.view.height == self.state.get_height():
                msgs = self.messages.get_valid_messages(m.view, m.type, lambda _: True)
                if self.has_quorum_by_msg_type(msgs, m.type):
                    self.messages.signal_event(m.type, m.view)
