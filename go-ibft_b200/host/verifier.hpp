// verifier.hpp -- the plugin boundary of the hot path, mirrored from reference core/backend.go:37-56 (core.Verifier) plus
// Backend.ID (:84), and its GPU-backed implementation.
//
//   Verifier        same five methods, same argument meaning, bool-only results ("never panic, malformed => false":
//                   SURVEY.md §8b).  Prefetch() is the one addition: the batching shim announces the messages a handler is
//                   about to validate so that a GPU backend can verify all their signatures in ONE device call; the
//                   per-message methods then answer from the verdict cache, which keeps the reference's serial predicate
//                   code (and therefore its exact semantics, including pruning order) untouched.
//   CallbackVerifier  closure-backed verifier = the reference's mockBackend (core/mock_test.go:72-151): defaults true,
//                   IsProposer default false.  Used by the CPU tests of the host logic.
//   GpuVerifier     the production path: every signature check goes through the C ABI (include/ibft_verify.h) into the
//                   CUDA kernels.  There is NO CPU fallback: if the engine call fails the answer is `false` (no verdict is
//                   ever invented) and last_error() says why.
#pragma once
#include <functional>
#include <mutex>
#include <unordered_map>

#include "../../include/ibft_verify.h"
#include "quorum.hpp"
#include "store.hpp"

namespace ibft::host {

class Verifier {
 public:
  virtual ~Verifier() = default;
  virtual bool IsValidProposal(const Bytes& raw_proposal) = 0;                                   // backend.go:39
  virtual bool IsValidValidator(const IbftMessage& msg) = 0;                                     // backend.go:41-45
  virtual bool IsProposer(const Bytes& id, uint64_t height, uint64_t round) = 0;                 // backend.go:47-48
  virtual bool IsValidProposalHash(const Proposal* proposal, const Bytes* hash) = 0;             // backend.go:50-51
  virtual bool IsValidCommittedSeal(const Bytes* proposal_hash, const CommittedSeal* seal) = 0;  // backend.go:53-55
  virtual Bytes ID() = 0;                                                                        // backend.go:84
  // Batching hook (no counterpart in the reference): sender signatures of `msgs` (+ their committed seals when
  // `with_seals`) are about to be checked.  Default: nothing.
  virtual void Prefetch(const std::vector<MessagePtr>& msgs, bool with_seals) {
    (void)msgs;
    (void)with_seals;
  }
};

class CallbackVerifier : public Verifier {
 public:
  std::function<bool(const Bytes&)> isValidProposalFn;
  std::function<bool(const IbftMessage&)> isValidValidatorFn;
  std::function<bool(const Bytes&, uint64_t, uint64_t)> isProposerFn;
  std::function<bool(const Proposal*, const Bytes*)> isValidProposalHashFn;
  std::function<bool(const Bytes*, const CommittedSeal*)> isValidCommittedSealFn;
  Bytes id;
  bool IsValidProposal(const Bytes& raw) override { return isValidProposalFn ? isValidProposalFn(raw) : true; }
  bool IsValidValidator(const IbftMessage& m) override { return isValidValidatorFn ? isValidValidatorFn(m) : true; }
  bool IsProposer(const Bytes& i, uint64_t h, uint64_t r) override { return isProposerFn ? isProposerFn(i, h, r) : false; }
  bool IsValidProposalHash(const Proposal* p, const Bytes* h) override { return isValidProposalHashFn ? isValidProposalHashFn(p, h) : true; }
  bool IsValidCommittedSeal(const Bytes* h, const CommittedSeal* s) override { return isValidCommittedSealFn ? isValidCommittedSealFn(h, s) : true; }
  Bytes ID() override { return id; }
};

class GpuVerifier : public Verifier {
 public:
  // embedder policy that is not signature work: stays on the host
  std::function<bool(const Bytes&, uint64_t, uint64_t)> isProposerFn;
  std::function<bool(const Bytes&)> isValidProposalFn;
  Bytes id;
  // Submit PREPARE / COMMIT messages as RAW FRAMES (IBFT_KIND_WIRE*): the device derives PayloadNoSig, From and the
  // signature from the gossip frame itself, so the host never re-marshals.  Frames the device hands back
  // (IBFT_ITEM_NEEDS_HOST: non-canonical encoding) are re-submitted through the marshalled path.
  bool use_wire_frames = false;

  explicit GpuVerifier(const ibft_engine_params& params) {
    params_ = params;
    int rc = ibft_engine_create(&params_, &engine_);
    if (rc != IBFT_OK) {
      engine_ = nullptr;
      error_ = ibft_last_error();
    }
  }
  ~GpuVerifier() override {
    if (engine_) ibft_engine_destroy(engine_);
  }
  bool ok() const { return engine_ != nullptr; }
  const std::string& last_error() const { return error_; }
  ibft_engine* engine() { return engine_; }

  // ValidatorBackend.GetVotingPowers(height) pushed to the device (validator_manager.go:50-57): the table of `height`
  // occupies slot (height mod max_table_slots).  Addresses that are not 20 bytes can never equal a recovered signer and are
  // left out of the device table.
  bool SetValidators(uint64_t height, const std::vector<Bytes>& addrs, const std::vector<u320>& powers) {
    if (!engine_) return false;
    std::lock_guard<std::recursive_mutex> lk(mu_);
    std::vector<uint8_t> a, p;
    for (size_t i = 0; i < addrs.size(); i++) {
      if (addrs[i].size() != 20) continue;
      a.insert(a.end(), addrs[i].begin(), addrs[i].end());
      for (int k = 3; k >= 0; k--)
        for (int j = 7; j >= 0; j--) p.push_back((uint8_t)(powers[i].l[k] >> (8 * j)));
    }
    uint32_t slot = (uint32_t)(height % params_.max_table_slots);
    int rc = ibft_set_validators(engine_, slot, height, a.data(), p.data(), (uint32_t)(a.size() / 20));
    if (rc != IBFT_OK) {
      error_ = ibft_last_error();
      return false;
    }
    slot_height_[slot] = height;
    current_height_ = height;
    cache_.clear();
    return true;
  }
  // committed seals carry no height: they are checked against the validators of the running sequence
  void SetCurrentHeight(uint64_t h) {
    std::lock_guard<std::recursive_mutex> lk(mu_);
    current_height_ = h;
  }

  bool IsValidProposal(const Bytes& raw) override { return isValidProposalFn ? isValidProposalFn(raw) : true; }
  bool IsProposer(const Bytes& i, uint64_t h, uint64_t r) override { return isProposerFn ? isProposerFn(i, h, r) : false; }
  Bytes ID() override { return id; }

  // The reference calls the verifier concurrently (gossip goroutines through AddMessage, the round goroutine and two
  // watchers: core/ibft.go:335-347, :1128) while the store holds its per-type mutex: every entry point below is serialised
  // on one lock and never calls back into the store.
  bool IsValidValidator(const IbftMessage& m) override {
    std::lock_guard<std::recursive_mutex> lk(mu_);
    Pending p;
    if (!sender_item(m, p)) return false;
    return lookup_or_verify(p);
  }
  bool IsValidCommittedSeal(const Bytes* proposal_hash, const CommittedSeal* seal) override {
    std::lock_guard<std::recursive_mutex> lk(mu_);
    Pending p;
    if (!seal_item(proposal_hash, seal, p)) return false;
    return lookup_or_verify(p);
  }
  // Synthetic proposal-hash convention of SURVEY.md §8(c): Keccak-256(Keccak-256(rawProposal) || u64_be(round)); real
  // embedders hash an RLP header (out of scope).  Hashing runs on the device (ibft_keccak256_batch), once per proposal.
  bool IsValidProposalHash(const Proposal* proposal, const Bytes* hash) override {
    if (!engine_ || !proposal || !hash || hash->size() != 32) return false;
    std::lock_guard<std::recursive_mutex> lk(mu_);
    Bytes key = proposal->raw_proposal;
    for (int j = 7; j >= 0; j--) key.push_back((char)(proposal->round >> (8 * j)));
    auto it = hash_cache_.find(key);
    if (it == hash_cache_.end()) {
      uint8_t inner[32], outer[32];
      if (!keccak(proposal->raw_proposal, inner)) return false;
      Bytes second((const char*)inner, 32);
      second.append(key.end() - 8, key.end());
      if (!keccak(second, outer)) return false;
      it = hash_cache_.emplace(key, Bytes((const char*)outer, 32)).first;
    }
    return it->second == *hash;
  }

  void Prefetch(const std::vector<MessagePtr>& msgs, bool with_seals) override {
    std::lock_guard<std::recursive_mutex> lk(mu_);
    std::vector<Pending> batch;
    std::unordered_map<Bytes, size_t> seen;
    std::function<void(const IbftMessage&)> visit = [&](const IbftMessage& m) {
      Pending p;
      if (sender_item(m, p) && !cache_.count(p.key) && seen.emplace(p.key, batch.size()).second) batch.push_back(std::move(p));
      if (with_seals && m.payload_kind == PAYLOAD_COMMIT) {
        auto seal = ExtractCommittedSeal(m);
        Pending q;
        if (seal_item(ExtractCommitHash(m), seal.get(), q) && !cache_.count(q.key) && seen.emplace(q.key, batch.size()).second)
          batch.push_back(std::move(q));
      }
      // nested signatures: prepared certificates inside ROUND_CHANGE, round-change certificates inside PREPREPARE
      if (m.payload_kind == PAYLOAD_ROUND_CHANGE && m.round_change.latest_prepared_certificate) {
        auto& pc = *m.round_change.latest_prepared_certificate;
        if (pc.proposal_message) visit(*pc.proposal_message);
        for (auto& pm : pc.prepare_messages)
          if (pm) visit(*pm);
      }
      if (m.payload_kind == PAYLOAD_PREPREPARE && m.preprepare.certificate)
        for (auto& rc : m.preprepare.certificate->round_change_messages)
          if (rc) visit(*rc);
    };
    for (auto& m : msgs)
      if (m) visit(*m);
    verify_pending(batch);
  }

  // Re-verification of the >= Q committed seals of an imported block (Backend.InsertProposal, core/backend.go:78-81; the same
  // check a syncing node performs): all seals in ONE launch.  valid[i] = IsValidCommittedSeal(hash, seals[i]).
  std::vector<bool> VerifyCommittedSeals(const Bytes& proposal_hash, const std::vector<CommittedSeal>& seals) {
    std::lock_guard<std::recursive_mutex> lk(mu_);
    std::vector<Pending> batch;
    std::vector<Bytes> keys(seals.size());
    std::unordered_map<Bytes, size_t> seen;
    for (size_t i = 0; i < seals.size(); i++) {
      Pending p;
      if (!seal_item(&proposal_hash, &seals[i], p)) continue;
      keys[i] = p.key;
      if (!cache_.count(p.key) && seen.emplace(p.key, batch.size()).second) batch.push_back(std::move(p));
    }
    verify_pending(batch);
    std::vector<bool> valid(seals.size(), false);
    for (size_t i = 0; i < seals.size(); i++) {
      auto it = keys[i].empty() ? cache_.end() : cache_.find(keys[i]);
      valid[i] = it != cache_.end() && it->second;
    }
    return valid;
  }

  uint64_t device_calls() const { return device_calls_; }
  uint64_t items_verified() const { return items_verified_; }
  uint64_t frames_handed_back() const { return frames_handed_back_; }

 private:
  struct Pending {
    Bytes key;  // exact-bytes cache key
    ibft_sig_item item;
    Bytes payload;
    uint64_t height;
    const IbftMessage* fallback_payload_msg = nullptr;  // raw-frame items: the message to re-marshal if the device hands it back
  };
  std::recursive_mutex mu_;
  ibft_engine_params params_{};
  ibft_engine* engine_ = nullptr;
  std::string error_;
  std::unordered_map<Bytes, bool> cache_;
  std::unordered_map<Bytes, Bytes> hash_cache_;
  std::map<uint32_t, uint64_t> slot_height_;
  uint64_t current_height_ = 0;
  uint64_t device_calls_ = 0, items_verified_ = 0, frames_handed_back_ = 0;

  bool keccak(const Bytes& data, uint8_t out[32]) {
    uint32_t off = 0, len = (uint32_t)data.size();
    int rc = ibft_keccak256_batch(engine_, (const uint8_t*)data.data(), data.size(), &off, &len, 1, out);
    device_calls_++;
    if (rc != IBFT_OK) error_ = ibft_last_error();
    return rc == IBFT_OK;
  }
  static void put_sig(ibft_sig_item& it, const Bytes& sig, const Bytes& signer) {
    memset(&it, 0, sizeof it);
    memcpy(it.r, sig.data(), 32);
    memcpy(it.s, sig.data() + 32, 32);
    it.v = (uint8_t)sig[64];
    memcpy(it.signer, signer.data(), 20);
  }
  // IsValidValidator: signer of msg.Signature over Keccak-256(PayloadNoSig) == msg.From and From is a validator at
  // msg.View.Height (backend.go:41-45).  Structurally invalid => false without touching the device.
  bool sender_item(const IbftMessage& m, Pending& p) {
    if (!m.view || m.from.size() != 20 || m.signature.size() != 65) return false;
    p.height = m.view->height;
    if (use_wire_frames && !m.raw_wire.empty() && (m.payload_kind == PAYLOAD_PREPARE || m.payload_kind == PAYLOAD_COMMIT)) {
      memset(&p.item, 0, sizeof p.item);
      p.item.kind = IBFT_KIND_WIRE;
      p.payload = m.raw_wire;
      p.fallback_payload_msg = &m;
      p.key.assign(1, 'W');
      for (int j = 7; j >= 0; j--) p.key.push_back((char)(p.height >> (8 * j)));
      p.key += m.raw_wire;
      return true;
    }
    p.payload = payload_no_sig(m);
    put_sig(p.item, m.signature, m.from);
    p.item.kind = IBFT_KIND_PAYLOAD;
    p.key.assign(1, 'S');
    for (int j = 7; j >= 0; j--) p.key.push_back((char)(p.height >> (8 * j)));
    p.key += m.signature;
    p.key += p.payload;
    return true;
  }
  bool seal_item(const Bytes* proposal_hash, const CommittedSeal* seal, Pending& p) {
    if (!proposal_hash || !seal || proposal_hash->size() != 32 || seal->signer.size() != 20 || seal->signature.size() != 65) return false;
    p.height = current_height_;
    put_sig(p.item, seal->signature, seal->signer);
    memcpy(p.item.digest, proposal_hash->data(), 32);
    p.item.kind = IBFT_KIND_SEAL;
    p.key.assign(1, 'C');
    for (int j = 7; j >= 0; j--) p.key.push_back((char)(p.height >> (8 * j)));
    p.key += seal->signature;
    p.key += seal->signer;
    p.key += *proposal_hash;
    return true;
  }
  bool lookup_or_verify(Pending& p) {
    auto it = cache_.find(p.key);
    if (it != cache_.end()) return it->second;
    std::vector<Pending> one;
    one.push_back(std::move(p));
    Bytes key = one[0].key;
    verify_pending(one);
    it = cache_.find(key);
    return it != cache_.end() && it->second;
  }
  // one device call for the whole batch; groups = distinct heights (validator tables)
  void verify_pending(std::vector<Pending>& batch) {
    if (batch.empty() || !engine_) return;
    size_t pos = 0;
    while (pos < batch.size()) {  // respect the engine's per-call capacity
      size_t n = std::min(batch.size() - pos, (size_t)params_.max_items);
      std::vector<ibft_sig_item> items(n);
      std::vector<ibft_group_desc> groups;
      std::map<uint64_t, uint16_t> group_of_height;
      Bytes arena;
      for (size_t i = 0; i < n; i++) {
        Pending& p = batch[pos + i];
        auto g = group_of_height.find(p.height);
        if (g == group_of_height.end()) {
          if (groups.size() >= params_.max_groups) { n = i; break; }
          uint32_t slot = (uint32_t)(p.height % params_.max_table_slots);
          auto sh = slot_height_.find(slot);
          ibft_group_desc d{};
          // a height whose validator table is not resident cannot have members: its items are verified against an
          // empty answer (false) -- never against the wrong table
          d.table_slot = (sh != slot_height_.end() && sh->second == p.height) ? (uint16_t)slot : (uint16_t)IBFT_NO_TABLE;
          groups.push_back(d);
          g = group_of_height.emplace(p.height, (uint16_t)(groups.size() - 1)).first;
        }
        items[i] = p.item;
        items[i].group = g->second;
        if (p.item.kind == IBFT_KIND_PAYLOAD || p.item.kind == IBFT_KIND_WIRE) {
          if (arena.size() + p.payload.size() > params_.max_payload_bytes) { n = i; break; }
          items[i].payload_off = (uint32_t)arena.size();
          items[i].payload_len = (uint32_t)p.payload.size();
          arena += p.payload;
        }
      }
      if (n == 0) {  // a single item that cannot fit: no verdict
        pos++;
        continue;
      }
      std::vector<uint32_t> bitmap((n + 31) / 32);
      int rc = ibft_verify_batch(engine_, items.data(), (uint32_t)n, (const uint8_t*)arena.data(), arena.size(), groups.data(),
                                 (uint32_t)groups.size(), bitmap.data(), nullptr, nullptr);
      device_calls_++;
      if (rc != IBFT_OK) {
        error_ = ibft_last_error();  // launch failure => NO verdict is cached; callers see `false`
      } else {
        items_verified_ += n;
        std::vector<uint8_t> status(n, 0);
        if (use_wire_frames) ibft_last_item_status(engine_, status.data(), (uint32_t)n);
        std::vector<Pending> redo;
        for (size_t i = 0; i < n; i++) {
          bool pass = (bitmap[i >> 5] >> (i & 31)) & 1u;
          // membership requires a resident table for the item's height
          if (groups[items[i].group].table_slot == IBFT_NO_TABLE) pass = false;
          Pending& p = batch[pos + i];
          if (status[i] == IBFT_ITEM_NEEDS_HOST && p.fallback_payload_msg) {
            // the device declined the frame (not canonical): same check through the marshalled path, same cache key
            Pending q;
            bool saved = use_wire_frames;
            use_wire_frames = false;
            bool ok = sender_item(*p.fallback_payload_msg, q);
            use_wire_frames = saved;
            if (ok) { q.key = p.key; redo.push_back(std::move(q)); }
            else cache_[p.key] = false;
            continue;
          }
          cache_[p.key] = pass;
        }
        if (!redo.empty()) { frames_handed_back_ += redo.size(); verify_pending(redo); }
      }
      pos += n;
    }
  }
};

}  // namespace ibft::host
