"""Generate tests/golden/proto_wire.json from the REFERENCE's own embedded descriptor.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_proto_golden.py
It reads the serialized FileDescriptorProto that protoc-gen-go embedded in
/root/reference/messages/proto/messages.pb.go (var file_messages_proto_messages_proto_rawDesc,
line 688+), loads it into google.protobuf, serialises a fixed list of messages with it and
records (case name, python-model repr, wire hex, wire hex with signature cleared = PayloadNoSig per
messages/proto/helper.go:13-27).  tests/test_oracle_proto.py replays the cases against
oracle/ibft_proto.py, and the C++ host codec is checked against the same file.
"""
import json
import os
import re
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402

from oracle import ibft_proto as ip  # noqa: E402

REF = "/root/reference/messages/proto/messages.pb.go"


def load_reference_classes():
    src = open(REF).read()
    m = re.search(r"file_messages_proto_messages_proto_rawDesc = \[\]byte\{(.*?)\n\}", src, re.S)
    raw = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(1)))
    fdp = descriptor_pb2.FileDescriptorProto.FromString(raw)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName(n))  # noqa: E731
    return {n: get(n) for n in ("IbftMessage", "View", "Proposal", "PreparedCertificate", "RoundChangeCertificate")}


def to_pb(cls, m: ip.IbftMessage):
    pb = cls["IbftMessage"]()
    fill_pb(pb, m)
    return pb


def fill_pb(pb, m: ip.IbftMessage):
    if m.view is not None:
        pb.view.SetInParent()
        pb.view.height = m.view.height
        pb.view.round = m.view.round
    pb.__setattr__("from", m.from_) if False else setattr(pb, "from", m.from_)
    pb.signature = m.signature
    pb.type = m.type
    p = m.payload
    if isinstance(p, ip.PrePrepareMessage):
        pb.preprepareData.SetInParent()
        if p.proposal is not None:
            pb.preprepareData.proposal.SetInParent()
            pb.preprepareData.proposal.rawProposal = p.proposal.raw_proposal
            pb.preprepareData.proposal.round = p.proposal.round
        pb.preprepareData.proposalHash = p.proposal_hash
        if p.certificate is not None:
            pb.preprepareData.certificate.SetInParent()
            for rc in p.certificate.round_change_messages:
                fill_pb(pb.preprepareData.certificate.roundChangeMessages.add(), rc)
    elif isinstance(p, ip.PrepareMessage):
        pb.prepareData.SetInParent()
        pb.prepareData.proposalHash = p.proposal_hash
    elif isinstance(p, ip.CommitMessage):
        pb.commitData.SetInParent()
        pb.commitData.proposalHash = p.proposal_hash
        pb.commitData.committedSeal = p.committed_seal
    elif isinstance(p, ip.RoundChangeMessage):
        pb.roundChangeData.SetInParent()
        if p.last_prepared_proposal is not None:
            pb.roundChangeData.lastPreparedProposal.SetInParent()
            pb.roundChangeData.lastPreparedProposal.rawProposal = p.last_prepared_proposal.raw_proposal
            pb.roundChangeData.lastPreparedProposal.round = p.last_prepared_proposal.round
        pc = p.latest_prepared_certificate
        if pc is not None:
            pb.roundChangeData.latestPreparedCertificate.SetInParent()
            if pc.proposal_message is not None:
                fill_pb(pb.roundChangeData.latestPreparedCertificate.proposalMessage, pc.proposal_message)
                pb.roundChangeData.latestPreparedCertificate.proposalMessage.SetInParent()
            for pm in pc.prepare_messages or []:
                fill_pb(pb.roundChangeData.latestPreparedCertificate.prepareMessages.add(), pm)


def cases():
    A = lambda i: bytes([0xA0 + i]) * 20  # noqa: E731
    H = b"\x22" * 32
    S = bytes(range(65))
    out = {}
    out["prepare_h1_r0"] = ip.IbftMessage(ip.View(1, 0), bytes(range(0xA0, 0xB4)), b"", ip.PREPARE, ip.PrepareMessage(H))
    out["prepare_h1e6_signed"] = ip.IbftMessage(ip.View(10**6, 0), A(1), S, ip.PREPARE, ip.PrepareMessage(H))
    out["commit_h1e6_signed"] = ip.IbftMessage(ip.View(10**6, 3), A(2), S, ip.COMMIT, ip.CommitMessage(H, S[::-1]))
    out["commit_empty_payload"] = ip.IbftMessage(ip.View(5, 0), A(2), S, ip.COMMIT, ip.CommitMessage())
    out["nil_view"] = ip.IbftMessage(None, A(3), S, ip.PREPARE, ip.PrepareMessage(H))
    out["empty_view"] = ip.IbftMessage(ip.View(0, 0), A(3), S, ip.PREPARE, ip.PrepareMessage(H))
    out["no_payload"] = ip.IbftMessage(ip.View(7, 1), A(4), S, ip.COMMIT, None)
    out["type_payload_mismatch"] = ip.IbftMessage(ip.View(7, 1), A(4), S, ip.PREPARE, ip.CommitMessage(H, S))
    out["all_default"] = ip.IbftMessage()
    out["big_varints"] = ip.IbftMessage(ip.View(2**64 - 1, 2**63), A(5), S, ip.ROUND_CHANGE, ip.RoundChangeMessage())
    pp0 = ip.IbftMessage(ip.View(9, 0), A(0), S, ip.PREPREPARE,
                         ip.PrePrepareMessage(ip.Proposal(b"block" * 50, 0), H, None))
    out["preprepare_r0"] = pp0
    preps = [ip.IbftMessage(ip.View(9, 0), A(i), S, ip.PREPARE, ip.PrepareMessage(H)) for i in (1, 2, 3)]
    pc = ip.PreparedCertificate(pp0, preps)
    rcs = [ip.IbftMessage(ip.View(9, 1), A(i), S, ip.ROUND_CHANGE,
                          ip.RoundChangeMessage(ip.Proposal(b"block" * 50, 0), pc)) for i in (0, 1, 2)]
    out["round_change_with_pc"] = rcs[0]
    out["round_change_empty"] = ip.IbftMessage(ip.View(9, 1), A(3), S, ip.ROUND_CHANGE, ip.RoundChangeMessage())
    out["round_change_pc_no_prepares"] = ip.IbftMessage(ip.View(9, 1), A(3), S, ip.ROUND_CHANGE,
                                                         ip.RoundChangeMessage(None, ip.PreparedCertificate(pp0, None)))
    out["round_change_pc_no_proposal"] = ip.IbftMessage(ip.View(9, 1), A(3), S, ip.ROUND_CHANGE,
                                                         ip.RoundChangeMessage(None, ip.PreparedCertificate(None, preps)))
    out["preprepare_r1_rcc"] = ip.IbftMessage(ip.View(9, 1), A(1), S, ip.PREPREPARE,
                                              ip.PrePrepareMessage(ip.Proposal(b"block" * 50, 1), H,
                                                                   ip.RoundChangeCertificate(rcs + [out["round_change_empty"]])))
    out["preprepare_empty_rcc"] = ip.IbftMessage(ip.View(9, 1), A(1), S, ip.PREPREPARE,
                                                 ip.PrePrepareMessage(ip.Proposal(b"", 0), b"", ip.RoundChangeCertificate([])))
    return out


def _tlvs(buf):
    out, pos = [], 0
    while pos < len(buf):
        s = pos
        key, pos = ip._read_varint(buf, pos)
        wt = key & 7
        if wt == 0:
            _, pos = ip._read_varint(buf, pos)
        elif wt == 2:
            ln, pos = ip._read_varint(buf, pos)
            pos += ln
        elif wt == 1:
            pos += 8
        elif wt == 5:
            pos += 4
        else:
            raise ValueError(wt)
        out.append(buf[s:pos])
    return out


def noncanonical_cases(cls):
    """Frames protobuf-go would NOT have produced but parses happily (a Byzantine validator may sign anything): duplicated and
    re-ordered fields, unknown fields, known numbers with the wrong wire type, groups, non-minimal varints, explicit zeros --
    with what google.protobuf (same descriptor) re-serialises them to, with and without the signature."""
    import random
    M = cls["IbftMessage"]
    rnd = random.Random(2024)
    unk = [bytes.fromhex(x) for x in ("4801", "5203616263", "5d01020304", "610102030405060708", "f80101", "1801", "0801", "2a00",
                                      "1a00", "7b08017c", "7b7c", "7b0a01617c", "2000", "20818000")]
    base = {n: to_pb(cls, m).SerializeToString(deterministic=True) for n, m in cases().items()}
    hand = {
        "view_split_in_two": bytes.fromhex("0a020805") + bytes.fromhex("0a021007") + base["prepare_h1_r0"][4:],
        "type_twice_last_wins": base["prepare_h1e6_signed"] + bytes.fromhex("2002"),
        "unknown_in_the_middle": base["commit_h1e6_signed"][:6] + bytes.fromhex("4801") + base["commit_h1e6_signed"][6:],
        "oneof_replaced": base["prepare_h1e6_signed"] + base["commit_h1e6_signed"][-103:],
        "signature_as_varint": bytes.fromhex("1801") + base["prepare_h1_r0"],
        "group_unknown": base["prepare_h1_r0"] + bytes.fromhex("7b0a0161087f7c"),
    }
    out = []
    for name, wire in hand.items():
        out.append((name, wire))
    names = sorted(base)
    for k in range(60):
        parts = _tlvs(base[names[k % len(names)]])
        for _ in range(rnd.randrange(1, 4)):
            op = rnd.randrange(5)
            if op == 0 and parts:
                parts.insert(rnd.randrange(len(parts) + 1), rnd.choice(parts))
            elif op == 1 and len(parts) > 1:
                i, j = rnd.randrange(len(parts)), rnd.randrange(len(parts))
                parts[i], parts[j] = parts[j], parts[i]
            elif op == 2:
                parts.insert(rnd.randrange(len(parts) + 1), rnd.choice(unk))
            elif op == 3 and parts:
                i = rnd.randrange(len(parts))
                key, pos = ip._read_varint(parts[i], 0)
                if key & 7 == 2:
                    ln, pos2 = ip._read_varint(parts[i], pos)
                    try:
                        inner = _tlvs(parts[i][pos2:pos2 + ln])
                    except Exception:
                        inner = []
                    if inner:
                        inner.insert(rnd.randrange(len(inner) + 1), rnd.choice(inner + unk))
                        body = b"".join(inner)
                        parts[i] = ip._varint(key) + ip._varint(len(body)) + body
            elif op == 4 and parts:
                del parts[rnd.randrange(len(parts))]
        out.append(("mutant_%02d" % k, b"".join(parts)))
    recs = []
    for name, wire in out:
        try:
            pb = M.FromString(wire)
        except Exception:
            recs.append({"name": name, "wire": wire.hex(), "parses": False})
            continue
        full = pb.SerializeToString(deterministic=True)
        pb.signature = b""
        recs.append({"name": name, "wire": wire.hex(), "parses": True, "remarshal": full.hex(),
                     "payload_no_sig": pb.SerializeToString(deterministic=True).hex()})
    return recs


def main():
    cls = load_reference_classes()
    recs = []
    for name, m in cases().items():
        pb = to_pb(cls, m)
        wire = pb.SerializeToString(deterministic=True)
        pb.signature = b""
        nosig = pb.SerializeToString(deterministic=True)
        recs.append({"name": name, "model": repr(m), "wire": wire.hex(), "payload_no_sig": nosig.hex()})
    path = os.path.join(os.path.dirname(__file__), "proto_wire.json")
    json.dump({"generator": "tests/golden/make_proto_golden.py",
               "source": "descriptor embedded in reference messages/proto/messages.pb.go (protoc-gen-go v1.28.1)",
               "cases": recs, "noncanonical": noncanonical_cases(cls)}, open(path, "w"), indent=1)
    print("wrote", path, len(recs), "cases")


if __name__ == "__main__":
    main()
