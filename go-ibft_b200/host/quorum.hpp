// quorum.hpp -- host-side mirror of go-ibft's ValidatorManager (reference core/validator_manager.go:23-155).
//
// The reference keeps voting power as math/big.Int; here it is a fixed 320-bit unsigned integer (the same width the device
// reduction uses: 256-bit powers summed over < 2^64 validators cannot overflow).  Besides the reference's address-set API
// (HasQuorum / HasPrepareQuorum) there is the variant the north star asks for: the quorum check READING THE GPU BITMAP --
// HasQuorumVoted(words) sums the power of the validators whose bit is set in the per-group voted set produced by the
// quorum kernels (ibft_get_voted_bitmap).
#pragma once
#include <cstring>
#include <map>
#include <set>
#include <mutex>
#include <shared_mutex>

#include "proto.hpp"

namespace ibft::host {

struct u320 {
  uint64_t l[5] = {0, 0, 0, 0, 0};
  static u320 from_be32(const uint8_t* b) {
    u320 r;
    for (int i = 0; i < 4; i++) {
      uint64_t w = 0;
      for (int j = 0; j < 8; j++) w = (w << 8) | b[(3 - i) * 8 + j];
      r.l[i] = w;
    }
    return r;
  }
  static u320 from_u64(uint64_t v) {
    u320 r;
    r.l[0] = v;
    return r;
  }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3] | l[4]) == 0; }
  void sub(const u320& o) {  // precondition: *this >= o
    unsigned __int128 b = 0;
    for (int i = 0; i < 5; i++) {
      unsigned __int128 t = (unsigned __int128)l[i] - o.l[i] - b;
      l[i] = (uint64_t)t;
      b = (t >> 64) & 1;
    }
  }
  void add(const u320& o) {
    unsigned __int128 c = 0;
    for (int i = 0; i < 5; i++) {
      c += (unsigned __int128)l[i] + o.l[i];
      l[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  int cmp(const u320& o) const {
    for (int i = 4; i >= 0; i--)
      if (l[i] != o.l[i]) return l[i] < o.l[i] ? -1 : 1;
    return 0;
  }
  // floor(2*this/3) + 1   (validator_manager.go:130-135)
  u320 quorum() const {
    uint64_t twice[6];
    uint64_t c = 0;
    for (int i = 0; i < 5; i++) {
      twice[i] = (l[i] << 1) | c;
      c = l[i] >> 63;
    }
    twice[5] = c;
    uint64_t q[6];
    unsigned __int128 rem = 0;
    for (int i = 5; i >= 0; i--) {
      unsigned __int128 cur = (rem << 64) | twice[i];
      q[i] = (uint64_t)(cur / 3);
      rem = cur % 3;
    }
    u320 r;
    unsigned __int128 k = 1;
    for (int i = 0; i < 5; i++) {
      k += q[i];
      r.l[i] = (uint64_t)k;
      k >>= 64;
    }
    return r;
  }
};

// stateType, core/state.go:10-18
enum StateName { NEW_ROUND = 0, PREPARE_STATE = 1, COMMIT_STATE = 2, FIN_STATE = 3 };

class ValidatorManager {
 public:
  // setCurrentVotingPower, validator_manager.go:61-74.  `order` fixes the validator index of every address (the order the
  // table was pushed to the engine with ibft_set_validators), so that voted-set bits can be mapped back.
  // Returns false for errVotingPowerNotCorrect (total <= 0).
  bool SetVotingPowers(const std::vector<Bytes>& order, const std::vector<u320>& powers) {
    std::unique_lock<std::shared_mutex> lk(mu_);
    u320 total;
    for (auto& p : powers) total.add(p);
    if (total.is_zero()) return false;
    power_.clear();
    order_ = order;
    powers_ = powers;
    for (size_t i = 0; i < order.size(); i++) power_[order[i]] = powers[i];
    quorum_ = total.quorum();
    initialised_ = true;
    epoch_++;
    return true;
  }
  // for the store's incremental per-view power accumulators: voting power of one address, and a counter that changes
  // whenever the table does (accumulators computed under another epoch are recomputed lazily)
  bool Lookup(const Bytes& addr, u320& out) const {
    std::shared_lock<std::shared_mutex> lk(mu_);
    auto it = power_.find(addr);
    if (it == power_.end()) return false;
    out = it->second;
    return true;
  }
  uint64_t epoch() const { return epoch_; }
  bool PowerReachesQuorum(const u320& power) const {
    std::shared_lock<std::shared_mutex> lk(mu_);
    return initialised_ && power.cmp(quorum_) >= 0;
  }
  bool initialised() const { return initialised_; }
  u320 quorum_size() const { return quorum_; }

  // HasQuorum, validator_manager.go:77-96: sum over the address SET, unknown addresses contribute nothing
  bool HasQuorum(const std::set<Bytes>& senders) const {
    std::shared_lock<std::shared_mutex> lk(mu_);
    if (!initialised_) return false;
    u320 power;
    for (auto& a : senders) {
      auto it = power_.find(a);
      if (it != power_.end()) power.add(it->second);
    }
    return power.cmp(quorum_) >= 0;
  }
  // The same test reading the GPU's voted set of a group: bit i <=> validator i (in `order`) has >= 1 valid message.
  bool HasQuorumVoted(const uint32_t* words, size_t n_words) const {
    std::shared_lock<std::shared_mutex> lk(mu_);
    if (!initialised_) return false;
    u320 power;
    for (size_t i = 0; i < powers_.size(); i++)
      if ((i >> 5) < n_words && ((words[i >> 5] >> (i & 31)) & 1u)) power.add(powers_[i]);
    return power.cmp(quorum_) >= 0;
  }
  // HasPrepareQuorum, validator_manager.go:99-127
  bool HasPrepareQuorum(StateName state, const MessagePtr& proposal_message, const std::vector<MessagePtr>& msgs,
                        std::vector<std::string>* errors = nullptr) const {
    if (!proposal_message) {
      if (state == PREPARE_STATE && errors) errors->push_back("HasPrepareQuorum - proposalMessage is not set");
      return false;
    }
    const Bytes& proposer = proposal_message->from;
    std::set<Bytes> senders{proposer};
    for (auto& m : msgs) {
      if (m->from == proposer) {
        if (errors) errors->push_back("HasPrepareQuorum - proposer is among signers but it is not expected to be");
        return false;
      }
      senders.insert(m->from);
    }
    return HasQuorum(senders);
  }

 private:
  mutable std::shared_mutex mu_;  // vpLock
  bool initialised_ = false;
  uint64_t epoch_ = 0;
  std::map<Bytes, u320> power_;
  std::vector<Bytes> order_;
  std::vector<u320> powers_;
  u320 quorum_;
};

inline std::set<Bytes> convertMessageToAddressSet(const std::vector<MessagePtr>& msgs) {  // validator_manager.go:147-155
  std::set<Bytes> s;
  for (auto& m : msgs) s.insert(m->from);
  return s;
}

}  // namespace ibft::host
