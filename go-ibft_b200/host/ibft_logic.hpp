// ibft_logic.hpp -- host-side mirror of the VALIDATION half of reference core/ibft.go (the callers of the hot path):
//   AddMessage / isAcceptableMessage          core/ibft.go:1101-1149
//   handlePrePrepare, validateProposal*       core/ibft.go:792-813, 629-788
//   handlePrepare                             core/ibft.go:855-889
//   handleCommit                              core/ibft.go:931-967
//   handleRoundChangeMessage                  core/ibft.go:470-512
//   proposalMatchesCertificate                core/ibft.go:516-551
//   validPC                                   core/ibft.go:1162-1231
//   hasQuorumByMsgType                        core/ibft.go:1273-1284
// The control half (round loop, timers, goroutines, transport) is out of scope and stays Go.
//
// The predicate code below is the reference's, statement for statement.  Batching is added only in front of it: when
// `batching` is on, each handler first hands the stored messages of the view to Verifier::Prefetch -- one GPU launch for
// all sender signatures, committed seals and nested certificate signatures -- and then runs the unchanged serial closures,
// which now answer from the verdict cache.  Decisions, pruning and returned sets are therefore identical by construction
// and are checked against oracle/ibft_logic.py in tests/test_host_logic.py.
#pragma once
#include "quorum.hpp"
#include "store.hpp"
#include "verifier.hpp"

namespace ibft::host {

struct State {  // the fields of core/state.go the predicates read or write
  View view;
  MessagePtr proposal_message;
  StateName name = NEW_ROUND;
  std::shared_ptr<PreparedCertificate> latest_pc;
  std::shared_ptr<Proposal> latest_prepared_proposal;
  std::vector<CommittedSeal> seals;
  const Proposal* getProposal() const {  // state.go:135-144
    return proposal_message ? ExtractProposal(*proposal_message) : nullptr;
  }
};

class IBFT {
 public:
  Verifier& backend;
  ValidatorManager& validatorManager;
  Messages& messages;
  State state;
  bool batching = true;
  bool incremental_quorum = false;  // O(log N) quorum probe in AddMessage instead of the reference's O(N) recomputation
  uint64_t commits_sent = 0;
  std::vector<std::string> log_errors;

  IBFT(Verifier& b, ValidatorManager& vm, Messages& ms) : backend(b), validatorManager(vm), messages(ms) { messages.SetPowerSource(&vm); }

  // core/ibft.go:1273-1284
  bool hasQuorumByMsgType(const std::vector<MessagePtr>& msgs, uint32_t type) {
    switch (type) {
      case PREPREPARE: return msgs.size() >= 1;
      case PREPARE: return validatorManager.HasPrepareQuorum(state.name, state.proposal_message, msgs, &log_errors);
      case ROUND_CHANGE:
      case COMMIT: return validatorManager.HasQuorum(convertMessageToAddressSet(msgs));
      default: return false;
    }
  }

  // core/ibft.go:1162-1231
  bool validPC(const PreparedCertificate* certificate, uint64_t roundLimit, uint64_t height) {
    if (!certificate) return true;  // PCs that are not set are valid by default
    if (!certificate->proposal_message || certificate->prepare_messages.empty()) return false;  // nil slice == empty on the wire
    std::vector<MessagePtr> allMessages{certificate->proposal_message};
    allMessages.insert(allMessages.end(), certificate->prepare_messages.begin(), certificate->prepare_messages.end());
    if (!validatorManager.HasQuorum(convertMessageToAddressSet(allMessages))) return false;
    if (certificate->proposal_message->type != PREPREPARE) return false;
    for (auto& m : certificate->prepare_messages)
      if (m->type != PREPARE) return false;
    if (!AreValidPCMessages(allMessages, height, roundLimit)) return false;
    const IbftMessage& proposal = *certificate->proposal_message;
    if (!backend.IsProposer(proposal.from, proposal.view->height, proposal.view->round)) return false;
    if (!backend.IsValidValidator(proposal)) return false;
    for (auto& m : certificate->prepare_messages) {
      if (!backend.IsValidValidator(*m)) return false;
      if (backend.IsProposer(m->from, m->view->height, m->view->round)) return false;
    }
    return true;
  }

  // core/ibft.go:516-551
  bool proposalMatchesCertificate(const Proposal* proposal, const PreparedCertificate* certificate) {
    if (!proposal && !certificate) return true;
    if (!certificate) return false;
    std::vector<const Bytes*> hashes;
    hashes.push_back(certificate->proposal_message ? ExtractProposalHash(*certificate->proposal_message) : nullptr);
    for (auto& m : certificate->prepare_messages) hashes.push_back(ExtractPrepareHash(*m));
    for (auto* h : hashes)
      if (!backend.IsValidProposalHash(proposal, h)) return false;
    return true;
  }

  // core/ibft.go:629-655
  bool validateProposalCommon(const IbftMessage& msg, const View& view) {
    const Proposal* proposal = ExtractProposal(msg);
    const Bytes* proposalHash = ExtractProposalHash(msg);
    if (!proposal) return false;  // the Go code would dereference nil; "never panic" => false
    if (proposal->round != view.round) return false;
    if (!backend.IsProposer(msg.from, view.height, view.round)) return false;
    if (!backend.IsValidProposalHash(proposal, proposalHash)) return false;
    return backend.IsValidProposal(proposal->raw_proposal);
  }
  // core/ibft.go:658-680
  bool validateProposal0(const IbftMessage& msg, const View& view) {
    if (!msg.view || msg.view->round != 0) return false;
    if (!validateProposalCommon(msg, view)) return false;
    if (backend.IsProposer(backend.ID(), view.height, view.round)) return false;
    return true;
  }
  // core/ibft.go:683-788
  bool validateProposal(const IbftMessage& msg, const View& view) {
    if (!msg.view) return false;  // the Go code reads msg.View.Round below; "never panic" => false
    uint64_t height = view.height, round = view.round;
    const Proposal* proposal = ExtractProposal(msg);
    const RoundChangeCertificate* rcc = ExtractRoundChangeCertificate(msg);
    if (!validateProposalCommon(msg, view)) return false;
    if (!rcc) return false;
    if (!HasUniqueSenders(rcc->round_change_messages)) return false;
    if (!hasQuorumByMsgType(rcc->round_change_messages, ROUND_CHANGE)) return false;
    if (backend.IsProposer(backend.ID(), height, round)) return false;
    for (auto& rc : rcc->round_change_messages) {
      if (rc->type != ROUND_CHANGE) return false;
      if (!rc->view || rc->view->height != height) return false;
      if (rc->view->round != round) return false;
      if (!backend.IsValidValidator(*rc)) return false;
    }
    struct Tuple { uint64_t round; const Bytes* hash; };
    std::vector<Tuple> roundsAndPreparedBlockHashes;
    for (auto& rcMessage : rcc->round_change_messages) {
      const PreparedCertificate* cert = ExtractLatestPC(*rcMessage);
      if (cert && validPC(cert, msg.view->round, height))
        roundsAndPreparedBlockHashes.push_back({cert->proposal_message->view->round, ExtractProposalHash(*cert->proposal_message)});
    }
    if (roundsAndPreparedBlockHashes.empty()) return true;
    uint64_t maxRound = 0;
    const Bytes* expectedHash = nullptr;
    for (auto& t : roundsAndPreparedBlockHashes)
      if (t.round >= maxRound) { maxRound = t.round; expectedHash = t.hash; }
    Proposal p{proposal->raw_proposal, maxRound};
    return backend.IsValidProposalHash(&p, expectedHash);
  }

  // core/ibft.go:792-813
  MessagePtr handlePrePrepare(const View& view) {
    if (batching) backend.Prefetch(messages.Snapshot(view, PREPREPARE), false);
    auto isValidPrePrepare = [&](const MessagePtr& m) {
      if (view.round == 0) return validateProposal0(*m, view);
      return validateProposal(*m, view);
    };
    auto msgs = messages.GetValidMessages(view, PREPREPARE, isValidPrePrepare);
    return msgs.empty() ? nullptr : msgs[0];
  }

  // core/ibft.go:855-889
  bool handlePrepare(const View& view) {
    auto isValidPrepare = [&](const MessagePtr& m) { return backend.IsValidProposalHash(state.getProposal(), ExtractPrepareHash(*m)); };
    auto prepareMessages = messages.GetValidMessages(view, PREPARE, isValidPrepare);
    if (!hasQuorumByMsgType(prepareMessages, PREPARE)) return false;
    commits_sent++;  // sendCommitMessage(view): transport is the embedder's
    auto pc = std::make_shared<PreparedCertificate>();
    pc->proposal_message = state.proposal_message;
    pc->prepare_messages = prepareMessages;
    state.latest_pc = pc;  // state.finalizePrepare
    const Proposal* p = state.getProposal();
    state.latest_prepared_proposal = p ? std::make_shared<Proposal>(*p) : nullptr;
    state.name = COMMIT_STATE;
    return true;
  }

  // core/ibft.go:931-967
  bool handleCommit(const View& view) {
    if (batching) backend.Prefetch(messages.Snapshot(view, COMMIT), true);  // all committed seals of the view: one launch
    auto isValidCommit = [&](const MessagePtr& m) {
      const Bytes* proposalHash = ExtractCommitHash(*m);
      auto committedSeal = ExtractCommittedSeal(*m);
      if (!backend.IsValidProposalHash(state.getProposal(), proposalHash)) return false;
      return backend.IsValidCommittedSeal(proposalHash, committedSeal.get());
    };
    auto commitMessages = messages.GetValidMessages(view, COMMIT, isValidCommit);
    if (!hasQuorumByMsgType(commitMessages, COMMIT)) return false;
    std::vector<CommittedSeal> seals;
    if (!ExtractCommittedSeals(commitMessages, seals)) {
      log_errors.push_back("failed to extract committed seals from commit messages");
      return false;
    }
    state.seals = seals;
    state.name = FIN_STATE;
    return true;
  }

  // core/ibft.go:470-512; *found=false corresponds to the nil certificate
  std::vector<MessagePtr> handleRoundChangeMessage(const View& view, bool* found) {
    uint64_t height = view.height;
    bool hasAcceptedProposal = state.getProposal() != nullptr;
    if (batching) backend.Prefetch(messages.SnapshotHeight(height, ROUND_CHANGE), false);  // nested PC signatures included
    auto isValidMsgFn = [&](const MessagePtr& m) {
      const Proposal* proposal = ExtractLastPreparedProposal(*m);
      const PreparedCertificate* certificate = ExtractLatestPC(*m);
      if (!validPC(certificate, m->view->round, height)) return false;
      return proposalMatchesCertificate(proposal, certificate);
    };
    auto isValidRCCFn = [&](uint64_t round, const std::vector<MessagePtr>& msgs) {
      if (round == view.round && hasAcceptedProposal) return false;
      return hasQuorumByMsgType(msgs, ROUND_CHANGE);
    };
    return messages.GetExtendedRCC(height, isValidMsgFn, isValidRCCFn, found);
  }

  // core/ibft.go:1126-1149
  bool isAcceptableMessage(const IbftMessage& message) {
    if (!backend.IsValidValidator(message)) return false;
    if (!message.view) return false;
    if (state.view.height > message.view->height) return false;
    if (state.view.height == message.view->height) return message.view->round >= state.view.round;
    return true;
  }
  // core/ibft.go:1101-1123
  void AddMessage(const MessagePtr& message) {
    if (!message) return;
    if (isAcceptableMessage(*message)) {
      messages.AddMessage(message);
      if (message->view->height == state.view.height) {
        bool quorum;
        if (!(incremental_quorum && quorumFromAccumulators(*message->view, message->type, &quorum))) {
          auto msgs = messages.GetValidMessages(*message->view, message->type, [](const MessagePtr&) { return true; });
          quorum = hasQuorumByMsgType(msgs, message->type);
        }
        if (quorum) messages.SignalEvent(message->type, *message->view);
      }
    }
  }
  // hasQuorumByMsgType (core/ibft.go:1273-1284) evaluated from the store's incremental accumulators: same decision as
  // building the address set and summing (validator_manager.go:77-127), without touching the N stored messages.
  bool quorumFromAccumulators(const View& view, uint32_t type, bool* quorum) {
    u320 power;
    size_t count = 0;
    bool proposer_among_senders = false;
    const Bytes* proposer = (type == PREPARE && state.proposal_message) ? &state.proposal_message->from : nullptr;
    if (!messages.ViewPower(view, type, &power, &count, proposer, &proposer_among_senders)) return false;
    switch (type) {
      case PREPREPARE: *quorum = count >= 1; return true;
      case PREPARE: {
        if (!state.proposal_message) {  // validator_manager.go:101-109
          if (state.name == PREPARE_STATE) log_errors.push_back("HasPrepareQuorum - proposalMessage is not set");
          *quorum = false;
          return true;
        }
        if (proposer_among_senders) {  // :116-121
          log_errors.push_back("HasPrepareQuorum - proposer is among signers but it is not expected to be");
          *quorum = false;
          return true;
        }
        u320 p;
        if (validatorManager.Lookup(*proposer, p)) power.add(p);
        *quorum = validatorManager.PowerReachesQuorum(power);
        return true;
      }
      case ROUND_CHANGE:
      case COMMIT: *quorum = validatorManager.PowerReachesQuorum(power); return true;
      default: *quorum = false; return true;
    }
  }

  // Bulk ingress: verify every inbound sender signature in one launch, then run the reference's
  // per-message AddMessage logic against the cache.
  void AddMessages(const std::vector<MessagePtr>& batch) {
    if (batching) backend.Prefetch(batch, false);
    for (auto& m : batch) AddMessage(m);
  }
};

}  // namespace ibft::host
