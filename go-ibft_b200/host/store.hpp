// store.hpp -- host-side mirror of go-ibft's message helpers and message store.
//   reference messages/helpers.go:16-227   CommittedSeal, Extract*, HasUniqueSenders, AreValidPCMessages
//   reference messages/messages.go:54-323  Messages: AddMessage (last write wins per sender), GetValidMessages (prunes
//                                          invalid messages), GetExtendedRCC (does NOT prune), GetMostRoundChangeMessages,
//                                          PruneByHeight, numMessages
// The event manager (messages/event_manager.go) is plumbing outside the hot path: only the SignalEvent
// hook is kept, as a callback.  Go iterates maps in unspecified order; here iteration is by sender key (deterministic).
#pragma once
#include <functional>
#include <map>
#include <mutex>
#include <set>

#include "proto.hpp"
#include "quorum.hpp"

namespace ibft::host {

// ------------------------------------------------------------------------------------------ messages/helpers.go
struct CommittedSeal {  // helpers.go:16-19
  Bytes signer, signature;
};

// helpers.go:38-48 -- nullptr when the payload is not CommitData (only the payload kind is checked there)
inline std::unique_ptr<CommittedSeal> ExtractCommittedSeal(const IbftMessage& m) {
  if (m.payload_kind != PAYLOAD_COMMIT) return nullptr;
  return std::make_unique<CommittedSeal>(CommittedSeal{m.from, m.commit.committed_seal});
}
// Each extractor returns nullptr on a type / payload mismatch (helpers.go:51-146); callers must treat that as "false".
inline const Bytes* ExtractCommitHash(const IbftMessage& m) {  // :51-62
  if (m.type != COMMIT || m.payload_kind != PAYLOAD_COMMIT) return nullptr;
  return &m.commit.proposal_hash;
}
inline const Proposal* ExtractProposal(const IbftMessage& m) {  // :65-76
  if (m.type != PREPREPARE || m.payload_kind != PAYLOAD_PREPREPARE) return nullptr;
  return m.preprepare.proposal.get();
}
inline const Bytes* ExtractProposalHash(const IbftMessage& m) {  // :79-90
  if (m.type != PREPREPARE || m.payload_kind != PAYLOAD_PREPREPARE) return nullptr;
  return &m.preprepare.proposal_hash;
}
inline const RoundChangeCertificate* ExtractRoundChangeCertificate(const IbftMessage& m) {  // :93-104
  if (m.type != PREPREPARE || m.payload_kind != PAYLOAD_PREPREPARE) return nullptr;
  return m.preprepare.certificate.get();
}
inline const Bytes* ExtractPrepareHash(const IbftMessage& m) {  // :107-118
  if (m.type != PREPARE || m.payload_kind != PAYLOAD_PREPARE) return nullptr;
  return &m.prepare.proposal_hash;
}
inline const PreparedCertificate* ExtractLatestPC(const IbftMessage& m) {  // :121-132
  if (m.type != ROUND_CHANGE || m.payload_kind != PAYLOAD_ROUND_CHANGE) return nullptr;
  return m.round_change.latest_prepared_certificate.get();
}
inline const Proposal* ExtractLastPreparedProposal(const IbftMessage& m) {  // :135-146
  if (m.type != ROUND_CHANGE || m.payload_kind != PAYLOAD_ROUND_CHANGE) return nullptr;
  return m.round_change.last_prepared_proposal.get();
}

// helpers.go:22-35: false (error) when a non-COMMIT message is present
inline bool ExtractCommittedSeals(const std::vector<MessagePtr>& msgs, std::vector<CommittedSeal>& out) {
  out.clear();
  for (const auto& m : msgs) {
    if (m->type != COMMIT) return false;
    auto s = ExtractCommittedSeal(*m);
    out.push_back(s ? *s : CommittedSeal{});
  }
  return true;
}

inline bool HasUniqueSenders(const std::vector<MessagePtr>& msgs) {  // helpers.go:149-166
  if (msgs.empty()) return false;
  std::set<Bytes> seen;
  for (const auto& m : msgs)
    if (!seen.insert(m->from).second) return false;
  return true;
}

// helpers.go:169-227.  Precondition (as in the reference, which would panic): every message has a non-nil View.
inline bool AreValidPCMessages(const std::vector<MessagePtr>& msgs, uint64_t height, uint64_t round_limit) {
  if (msgs.empty()) return false;
  if (!msgs[0]->view) return false;
  uint64_t round = msgs[0]->view->round;
  std::set<Bytes> senders;
  const Bytes* hash = nullptr;
  bool have_hash = false;
  for (const auto& m : msgs) {
    if (!m->view) return false;  // the Go code would dereference nil here; "never panic" => false
    if (m->view->height != height) return false;
    if (m->view->round != round || m->view->round >= round_limit) return false;
    const Bytes* extracted = nullptr;
    bool ok = false;
    if (m->type == PREPREPARE) { extracted = ExtractProposalHash(*m); ok = true; }
    else if (m->type == PREPARE) { extracted = ExtractPrepareHash(*m); ok = true; }
    // `if hash == nil { hash = extractedHash }`: a wire-decoded empty hash is a nil slice and never becomes the reference
    if (!have_hash && extracted && !extracted->empty()) { hash = extracted; have_hash = true; }
    static const Bytes kEmpty;
    const Bytes& h = hash ? *hash : kEmpty;
    const Bytes& e = extracted ? *extracted : kEmpty;
    if (!ok || h != e) return false;  // bytes.Equal: nil == empty
    if (!senders.insert(m->from).second) return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------------ messages/messages.go
using IsValidFn = std::function<bool(const MessagePtr&)>;
using IsValidRCCFn = std::function<bool(uint64_t round, const std::vector<MessagePtr>&)>;

class Messages {
 public:
  using SignalFn = std::function<void(uint32_t type, uint64_t height, uint64_t round)>;
  SignalFn on_signal;  // stands in for the event manager (messages.go:68-72)

  // Incremental quorum bookkeeping.  The reference recomputes, for EVERY inbound message, a copy of all
  // N stored messages of the view, an N-entry address set and an N-term big.Int sum (core/ibft.go:1113-1120,
  // validator_manager.go:77-96, :147-155): O(N^2) per round.  With a power source attached, every view bucket keeps the
  // summed voting power of its distinct senders up to date on insert / prune, and ViewPower answers in O(log N).
  void SetPowerSource(const ValidatorManager* vm) { vm_ = vm; }

  void AddMessage(const MessagePtr& m) {  // :54-65
    if (!m || !m->view || m->type > ROUND_CHANGE) return;
    std::lock_guard<std::mutex> lk(mux_[m->type]);
    ViewBucket& b = maps_[m->type][m->view->height][m->view->round];
    auto ins = b.msgs.insert_or_assign(m->from, m);
    if (ins.second && vm_ && b.epoch == vm_->epoch()) {  // a NEW sender (a replacement leaves the sender set unchanged)
      u320 p;
      if (vm_->Lookup(m->from, p)) b.power.add(p);
    }
  }
  // summed voting power and number of the distinct senders stored for (view, type); has_sender tests one address.
  // Returns false when no power source is attached.
  bool ViewPower(const View& v, uint32_t type, u320* power, size_t* count, const Bytes* probe, bool* has_probe) {
    if (!vm_) return false;
    std::lock_guard<std::mutex> lk(mux_[type]);
    *power = u320();
    *count = 0;
    if (has_probe) *has_probe = false;
    ViewBucket* b = bucket(type, v);
    if (!b) return true;
    refresh(*b);
    *power = b->power;
    *count = b->msgs.size();
    if (probe && has_probe) *has_probe = b->msgs.count(*probe) != 0;
    return true;
  }
  void SignalEvent(uint32_t type, const View& v) {
    if (on_signal) on_signal(type, v.height, v.round);
  }
  size_t numMessages(const View& v, uint32_t type) {  // :96-119
    std::lock_guard<std::mutex> lk(mux_[type]);
    auto* msgs = find(type, v);
    return msgs ? msgs->size() : 0;
  }
  void PruneByHeight(uint64_t height) {  // :123-148
    for (uint32_t t = 0; t < 4; t++) {
      std::lock_guard<std::mutex> lk(mux_[t]);
      auto& hm = maps_[t];
      hm.erase(hm.begin(), hm.lower_bound(height));
    }
  }
  // :169-199 -- invalid messages are pruned out of the store
  std::vector<MessagePtr> GetValidMessages(const View& v, uint32_t type, const IsValidFn& is_valid) {
    std::lock_guard<std::mutex> lk(mux_[type]);
    std::vector<MessagePtr> valid;
    auto* msgs = find(type, v);
    if (!msgs) return valid;
    std::vector<Bytes> invalid_keys;
    for (auto& kv : *msgs) {
      if (!is_valid(kv.second)) { invalid_keys.push_back(kv.first); continue; }
      valid.push_back(kv.second);
    }
    if (!invalid_keys.empty()) {
      ViewBucket* b = bucket(type, v);
      for (auto& k : invalid_keys) {
        if (vm_ && b && b->epoch == vm_->epoch()) {
          u320 p;
          if (vm_->Lookup(k, p)) b->power.sub(p);
        }
        msgs->erase(k);
      }
    }
    return valid;
  }
  // Snapshot of the stored messages of a view (no validation, no pruning): what the batching shim feeds to the GPU
  // before GetValidMessages runs the per-message closure against the resulting verdict cache.
  std::vector<MessagePtr> Snapshot(const View& v, uint32_t type) {
    std::lock_guard<std::mutex> lk(mux_[type]);
    std::vector<MessagePtr> out;
    if (auto* msgs = find(type, v))
      for (auto& kv : *msgs) out.push_back(kv.second);
    return out;
  }
  std::vector<MessagePtr> SnapshotHeight(uint64_t height, uint32_t type) {
    std::lock_guard<std::mutex> lk(mux_[type]);
    std::vector<MessagePtr> out;
    auto it = maps_[type].find(height);
    if (it != maps_[type].end())
      for (auto& rm : it->second)
        for (auto& kv : rm.second.msgs) out.push_back(kv.second);
    return out;
  }
  // :202-245 -- highest round whose valid messages satisfy isValidRCC; does not prune.  `found` distinguishes the Go nil
  // result from an empty slice.
  std::vector<MessagePtr> GetExtendedRCC(uint64_t height, const IsValidFn& is_valid_msg, const IsValidRCCFn& is_valid_rcc,
                                         bool* found = nullptr) {
    std::lock_guard<std::mutex> lk(mux_[ROUND_CHANGE]);
    std::vector<MessagePtr> extended;
    bool have = false;
    uint64_t highest = 0;
    auto it = maps_[ROUND_CHANGE].find(height);
    if (it != maps_[ROUND_CHANGE].end()) {
      for (auto& rm : it->second) {
        uint64_t round = rm.first;
        if (round <= highest) continue;
        std::vector<MessagePtr> valid;
        for (auto& kv : rm.second.msgs)
          if (is_valid_msg(kv.second)) valid.push_back(kv.second);
        if (!is_valid_rcc(round, valid)) continue;
        highest = round;
        extended = valid;
        have = true;
      }
    }
    if (found) *found = have;
    return extended;
  }
  // :249-286
  std::vector<MessagePtr> GetMostRoundChangeMessages(uint64_t min_round, uint64_t height, bool* found = nullptr) {
    std::lock_guard<std::mutex> lk(mux_[ROUND_CHANGE]);
    std::vector<MessagePtr> out;
    uint64_t best_round = 0;
    size_t best = 0;
    auto it = maps_[ROUND_CHANGE].find(height);
    if (it != maps_[ROUND_CHANGE].end()) {
      for (auto& rm : it->second) {
        if (rm.first < min_round) continue;
        if (rm.second.msgs.size() > best) { best_round = rm.first; best = rm.second.msgs.size(); }
      }
      if (best_round != 0)
        for (auto& kv : it->second[best_round].msgs) out.push_back(kv.second);
    }
    if (found) *found = best_round != 0;
    return out;
  }

 private:
  using ProtoMessages = std::map<Bytes, MessagePtr>;            // sender -> message   (messages.go:296)
  struct ViewBucket {
    ProtoMessages msgs;
    u320 power;                         // sum of votingPower over the distinct senders that are validators
    uint64_t epoch = ~0ull;             // ValidatorManager epoch `power` was computed under (~0: never)
  };
  using RoundMap = std::map<uint64_t, ViewBucket>;               // round  -> messages  (:293)
  using HeightMap = std::map<uint64_t, RoundMap>;                // height -> rounds    (:290)
  HeightMap maps_[4];
  std::mutex mux_[4];  // one lock per message type (:44-49)
  const ValidatorManager* vm_ = nullptr;

  ViewBucket* bucket(uint32_t type, const View& v) {
    auto h = maps_[type].find(v.height);
    if (h == maps_[type].end()) return nullptr;
    auto r = h->second.find(v.round);
    return r == h->second.end() ? nullptr : &r->second;
  }
  ProtoMessages* find(uint32_t type, const View& v) {
    ViewBucket* b = bucket(type, v);
    return b ? &b->msgs : nullptr;
  }
  void refresh(ViewBucket& b) {  // the validator table changed (new height): recompute once, O(N)
    if (!vm_ || b.epoch == vm_->epoch()) return;
    b.power = u320();
    for (auto& kv : b.msgs) {
      u320 p;
      if (vm_->Lookup(kv.first, p)) b.power.add(p);
    }
    b.epoch = vm_->epoch();
  }
};

}  // namespace ibft::host
