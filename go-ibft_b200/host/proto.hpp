// proto.hpp -- host-side mirror of go-ibft's wire schema (reference messages/proto/messages.proto:7-110) and of
// IbftMessage.PayloadNoSig (reference messages/proto/helper.go:13-27).
//
// The reference is Go; no Go toolchain exists in the build image, so the host side above the C ABI is C++ with the same
// names and semantics (INTEGRATION.md shows the Go/cgo binding a maintainer would add).  Encoding follows protobuf-go
// v1.28.1 for this schema: fields in field-number order, zero scalars / empty bytes omitted, nil sub-messages omitted,
// present-but-empty sub-messages as `tag 00`, the set oneof member always emitted.  Checked byte-for-byte against wire
// bytes produced from the reference's own descriptor (tests/golden/proto_wire.json).
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ibft::host {

using Bytes = std::string;  // arbitrary bytes; like Go's string(key) it is directly usable as a map key

enum MessageType : uint32_t { PREPREPARE = 0, PREPARE = 1, COMMIT = 2, ROUND_CHANGE = 3 };  // messages.proto:7-12

struct View {  // messages.proto:15-21
  uint64_t height = 0, round = 0;
};
struct Proposal {  // messages.proto:104-110
  Bytes raw_proposal;
  uint64_t round = 0;
};
struct IbftMessage;
using MessagePtr = std::shared_ptr<IbftMessage>;

struct PreparedCertificate {  // messages.proto:87-94
  MessagePtr proposal_message;                 // nullptr == Go nil
  std::vector<MessagePtr> prepare_messages;    // empty == Go nil slice (indistinguishable on the wire)
};
struct RoundChangeCertificate {  // messages.proto:98-101
  std::vector<MessagePtr> round_change_messages;
};
struct PrePrepareMessage {  // messages.proto:47-57
  std::shared_ptr<Proposal> proposal;
  Bytes proposal_hash;
  std::shared_ptr<RoundChangeCertificate> certificate;
};
struct PrepareMessage {  // messages.proto:60-63
  Bytes proposal_hash;
};
struct CommitMessage {  // messages.proto:66-72
  Bytes proposal_hash, committed_seal;
};
struct RoundChangeMessage {  // messages.proto:75-83
  std::shared_ptr<Proposal> last_prepared_proposal;
  std::shared_ptr<PreparedCertificate> latest_prepared_certificate;
};

enum PayloadKind : uint8_t { PAYLOAD_NONE = 0, PAYLOAD_PREPREPARE = 5, PAYLOAD_PREPARE = 6, PAYLOAD_COMMIT = 7, PAYLOAD_ROUND_CHANGE = 8 };

struct IbftMessage {  // messages.proto:24-44
  std::shared_ptr<View> view;  // nullptr == nil view
  Bytes from, signature;
  uint32_t type = PREPREPARE;
  PayloadKind payload_kind = PAYLOAD_NONE;  // which oneof member is set (independent of `type`, as in Go)
  // Where this message sits in the gossip frame it was decoded from: [wire_off, wire_off + wire_len) of *root_wire.  Set for the
  // top-level message AND for every message nested in its certificates (they share the root's bytes), so that the GPU verifier can
  // submit any of them as a raw-frame span without re-marshalling (IBFT_KIND_WIRE).  nullptr for messages built in memory.
  std::shared_ptr<const Bytes> root_wire;
  uint32_t wire_off = 0, wire_len = 0;
  bool has_wire() const { return root_wire != nullptr && wire_len != 0; }
  const char* wire_data() const { return root_wire->data() + wire_off; }
  PrePrepareMessage preprepare;
  PrepareMessage prepare;
  CommitMessage commit;
  RoundChangeMessage round_change;
};

struct DecodeError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ----------------------------------------------------------------------------------------------- encoder
namespace wire {
inline void varint(Bytes& o, uint64_t v) {
  while (v >= 0x80) {
    o.push_back((char)(v | 0x80));
    v >>= 7;
  }
  o.push_back((char)v);
}
inline void f_varint(Bytes& o, uint32_t num, uint64_t v) {
  if (v == 0) return;
  varint(o, (uint64_t)num << 3);
  varint(o, v);
}
inline void f_bytes(Bytes& o, uint32_t num, const Bytes& b) {
  if (b.empty()) return;
  varint(o, ((uint64_t)num << 3) | 2);
  varint(o, b.size());
  o += b;
}
inline void f_msg(Bytes& o, uint32_t num, const Bytes& body) {  // present (possibly empty) sub-message
  varint(o, ((uint64_t)num << 3) | 2);
  varint(o, body.size());
  o += body;
}
}  // namespace wire

Bytes encode_message(const IbftMessage& m, bool with_signature = true);

inline Bytes encode_proposal(const Proposal& p) {
  Bytes o;
  wire::f_bytes(o, 1, p.raw_proposal);
  wire::f_varint(o, 2, p.round);
  return o;
}
inline Bytes encode_pc(const PreparedCertificate& pc) {
  Bytes o;
  if (pc.proposal_message) wire::f_msg(o, 1, encode_message(*pc.proposal_message));
  for (const auto& m : pc.prepare_messages) wire::f_msg(o, 2, m ? encode_message(*m) : Bytes());
  return o;
}
inline Bytes encode_rcc(const RoundChangeCertificate& r) {
  Bytes o;
  for (const auto& m : r.round_change_messages) wire::f_msg(o, 1, m ? encode_message(*m) : Bytes());
  return o;
}

inline Bytes encode_message(const IbftMessage& m, bool with_signature) {
  Bytes o;
  if (m.view) {
    Bytes v;
    wire::f_varint(v, 1, m.view->height);
    wire::f_varint(v, 2, m.view->round);
    wire::f_msg(o, 1, v);
  }
  wire::f_bytes(o, 2, m.from);
  if (with_signature) wire::f_bytes(o, 3, m.signature);
  wire::f_varint(o, 4, m.type);
  switch (m.payload_kind) {
    case PAYLOAD_PREPREPARE: {
      Bytes b;
      if (m.preprepare.proposal) wire::f_msg(b, 1, encode_proposal(*m.preprepare.proposal));
      wire::f_bytes(b, 2, m.preprepare.proposal_hash);
      if (m.preprepare.certificate) wire::f_msg(b, 3, encode_rcc(*m.preprepare.certificate));
      wire::f_msg(o, 5, b);
      break;
    }
    case PAYLOAD_PREPARE: {
      Bytes b;
      wire::f_bytes(b, 1, m.prepare.proposal_hash);
      wire::f_msg(o, 6, b);
      break;
    }
    case PAYLOAD_COMMIT: {
      Bytes b;
      wire::f_bytes(b, 1, m.commit.proposal_hash);
      wire::f_bytes(b, 2, m.commit.committed_seal);
      wire::f_msg(o, 7, b);
      break;
    }
    case PAYLOAD_ROUND_CHANGE: {
      Bytes b;
      if (m.round_change.last_prepared_proposal) wire::f_msg(b, 1, encode_proposal(*m.round_change.last_prepared_proposal));
      if (m.round_change.latest_prepared_certificate) wire::f_msg(b, 2, encode_pc(*m.round_change.latest_prepared_certificate));
      wire::f_msg(o, 8, b);
      break;
    }
    default: break;
  }
  return o;
}

// messages/proto/helper.go:13-27: clone, Signature = nil, proto.Marshal -- the bytes that are hashed and signed
inline Bytes payload_no_sig(const IbftMessage& m) { return encode_message(m, false); }

// ----------------------------------------------------------------------------------------------- decoder
namespace wire {
// Nesting bound of the decoder.  A valid frame nests at most 4 messages deep (PREPREPARE -> RoundChangeCertificate ->
// ROUND_CHANGE -> PreparedCertificate -> PREPARE); decoding runs BEFORE any signature check, so an unauthenticated peer must
// not be able to exhaust the native stack with a few hundred KB of nested wrappers (protobuf-go caps recursion at 10,000 on
// growable stacks; a C++ frame is not growable).  Every sub-reader inherits depth + 1; decode_message refuses depth > kMaxDepth.
constexpr int kMaxDepth = 32;
struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  int depth = 0;
  const std::shared_ptr<const Bytes>* root = nullptr;  // the frame being decoded (sub-readers inherit it): messages record their span
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    for (;;) {
      if (p >= end || shift > 63) throw DecodeError("truncated/overlong varint");
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7F) << shift;
      shift += 7;
      if (!(b & 0x80)) return v;
    }
  }
  // returns false at end; otherwise field number, wire type and (for wt 0) value or (wt 2) the slice
  bool next(uint32_t& num, uint32_t& wt, uint64_t& val, Reader& sub) {
    if (done()) return false;
    uint64_t key = varint();
    num = (uint32_t)(key >> 3);
    wt = (uint32_t)(key & 7);
    if (num == 0) throw DecodeError("field number 0");
    switch (wt) {
      case 0: val = varint(); break;
      case 2: {
        uint64_t n = varint();
        if (n > (uint64_t)(end - p)) throw DecodeError("truncated bytes");
        sub = Reader{p, p + n, depth + 1, root};
        if (sub.depth > kMaxDepth) throw DecodeError("message nesting too deep");
        p += n;
        break;
      }
      case 1:
        if (end - p < 8) throw DecodeError("truncated fixed64");
        p += 8;
        break;
      case 5:
        if (end - p < 4) throw DecodeError("truncated fixed32");
        p += 4;
        break;
      default: throw DecodeError("unsupported wire type");
    }
    return true;
  }
  Bytes bytes() const { return Bytes((const char*)p, (size_t)(end - p)); }
};
}  // namespace wire

MessagePtr decode_message(wire::Reader r);

inline std::shared_ptr<Proposal> decode_proposal(wire::Reader r) {
  auto p = std::make_shared<Proposal>();
  uint32_t num, wt;
  uint64_t val;
  wire::Reader sub{nullptr, nullptr, 0, nullptr};
  while (r.next(num, wt, val, sub)) {
    if (num == 1 && wt == 2) p->raw_proposal = sub.bytes();
    else if (num == 2 && wt == 0) p->round = val;
  }
  return p;
}
inline std::shared_ptr<PreparedCertificate> decode_pc(wire::Reader r) {
  auto pc = std::make_shared<PreparedCertificate>();
  uint32_t num, wt;
  uint64_t val;
  wire::Reader sub{nullptr, nullptr, 0, nullptr};
  while (r.next(num, wt, val, sub)) {
    if (num == 1 && wt == 2) pc->proposal_message = decode_message(sub);
    else if (num == 2 && wt == 2) pc->prepare_messages.push_back(decode_message(sub));
  }
  return pc;
}
inline std::shared_ptr<RoundChangeCertificate> decode_rcc(wire::Reader r) {
  auto c = std::make_shared<RoundChangeCertificate>();
  uint32_t num, wt;
  uint64_t val;
  wire::Reader sub{nullptr, nullptr, 0, nullptr};
  while (r.next(num, wt, val, sub))
    if (num == 1 && wt == 2) c->round_change_messages.push_back(decode_message(sub));
  return c;
}

inline MessagePtr decode_message(wire::Reader r) {
  auto m = std::make_shared<IbftMessage>();
  if (r.root && *r.root) {
    m->root_wire = *r.root;
    m->wire_off = (uint32_t)(r.p - (const uint8_t*)(*r.root)->data());
    m->wire_len = (uint32_t)(r.end - r.p);
  }
  uint32_t num, wt;
  uint64_t val;
  wire::Reader sub{nullptr, nullptr, 0, nullptr};
  while (r.next(num, wt, val, sub)) {
    if (num == 1 && wt == 2) {
      auto v = std::make_shared<View>();
      uint32_t n2, w2;
      uint64_t v2;
      wire::Reader s2{nullptr, nullptr, 0, nullptr};
      while (sub.next(n2, w2, v2, s2)) {
        if (n2 == 1 && w2 == 0) v->height = v2;
        else if (n2 == 2 && w2 == 0) v->round = v2;
      }
      m->view = v;
    } else if (num == 2 && wt == 2) {
      m->from = sub.bytes();
    } else if (num == 3 && wt == 2) {
      m->signature = sub.bytes();
    } else if (num == 4 && wt == 0) {
      m->type = (uint32_t)val;
    } else if (num == 5 && wt == 2) {
      m->payload_kind = PAYLOAD_PREPREPARE;
      m->preprepare = PrePrepareMessage();
      uint32_t n2, w2;
      uint64_t v2;
      wire::Reader s2{nullptr, nullptr, 0, nullptr};
      while (sub.next(n2, w2, v2, s2)) {
        if (n2 == 1 && w2 == 2) m->preprepare.proposal = decode_proposal(s2);
        else if (n2 == 2 && w2 == 2) m->preprepare.proposal_hash = s2.bytes();
        else if (n2 == 3 && w2 == 2) m->preprepare.certificate = decode_rcc(s2);
      }
    } else if (num == 6 && wt == 2) {
      m->payload_kind = PAYLOAD_PREPARE;
      m->prepare = PrepareMessage();
      uint32_t n2, w2;
      uint64_t v2;
      wire::Reader s2{nullptr, nullptr, 0, nullptr};
      while (sub.next(n2, w2, v2, s2))
        if (n2 == 1 && w2 == 2) m->prepare.proposal_hash = s2.bytes();
    } else if (num == 7 && wt == 2) {
      m->payload_kind = PAYLOAD_COMMIT;
      m->commit = CommitMessage();
      uint32_t n2, w2;
      uint64_t v2;
      wire::Reader s2{nullptr, nullptr, 0, nullptr};
      while (sub.next(n2, w2, v2, s2)) {
        if (n2 == 1 && w2 == 2) m->commit.proposal_hash = s2.bytes();
        else if (n2 == 2 && w2 == 2) m->commit.committed_seal = s2.bytes();
      }
    } else if (num == 8 && wt == 2) {
      m->payload_kind = PAYLOAD_ROUND_CHANGE;
      m->round_change = RoundChangeMessage();
      uint32_t n2, w2;
      uint64_t v2;
      wire::Reader s2{nullptr, nullptr, 0, nullptr};
      while (sub.next(n2, w2, v2, s2)) {
        if (n2 == 1 && w2 == 2) m->round_change.last_prepared_proposal = decode_proposal(s2);
        else if (n2 == 2 && w2 == 2) m->round_change.latest_prepared_certificate = decode_pc(s2);
      }
    }
  }
  return m;
}

inline MessagePtr decode_message(const uint8_t* data, size_t len) {
  auto root = std::make_shared<const Bytes>((const char*)data, len);
  const uint8_t* b = (const uint8_t*)root->data();
  return decode_message(wire::Reader{b, b + len, 0, &root});
}

}  // namespace ibft::host
