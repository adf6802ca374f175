"""go-ibft wire schema (proto3) restated as plain Python: model, encoder, decoder, PayloadNoSig.

Oracle / test infrastructure only (see oracle/__init__.py).  Follows
messages/proto/messages.proto:7-110 (schema) and messages/proto/helper.go:13-27 (PayloadNoSig =
clone, Signature=nil, proto.Marshal).  Encoding rules are those of proto3 as emitted by
protobuf-go v1.28.1 (go.mod:9) for this schema: fields in field-number order, zero scalars and
empty bytes omitted, nil sub-messages omitted, present-but-empty sub-messages encoded as
``tag 00``, the set ``oneof payload`` member always emitted.  tests/test_oracle_proto.py checks
this encoder byte-for-byte against google.protobuf driven by the same schema.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Union

PREPREPARE, PREPARE, COMMIT, ROUND_CHANGE = 0, 1, 2, 3  # messages.proto:7-12


@dataclass
class View:  # messages.proto:15-21
    height: int = 0
    round: int = 0


@dataclass
class Proposal:  # messages.proto:104-110
    raw_proposal: bytes = b""
    round: int = 0


@dataclass
class PreparedCertificate:  # messages.proto:87-94
    proposal_message: Optional["IbftMessage"] = None
    prepare_messages: Optional[List["IbftMessage"]] = None  # None == Go nil slice


@dataclass
class RoundChangeCertificate:  # messages.proto:98-101
    round_change_messages: List["IbftMessage"] = field(default_factory=list)


@dataclass
class PrePrepareMessage:  # messages.proto:47-57
    proposal: Optional[Proposal] = None
    proposal_hash: bytes = b""
    certificate: Optional[RoundChangeCertificate] = None


@dataclass
class PrepareMessage:  # messages.proto:60-63
    proposal_hash: bytes = b""


@dataclass
class CommitMessage:  # messages.proto:66-72
    proposal_hash: bytes = b""
    committed_seal: bytes = b""


@dataclass
class RoundChangeMessage:  # messages.proto:75-83
    last_prepared_proposal: Optional[Proposal] = None
    latest_prepared_certificate: Optional[PreparedCertificate] = None


Payload = Union[PrePrepareMessage, PrepareMessage, CommitMessage, RoundChangeMessage, None]


@dataclass
class IbftMessage:  # messages.proto:24-44
    view: Optional[View] = None
    from_: bytes = b""
    signature: bytes = b""
    type: int = PREPREPARE
    payload: Payload = None

    def payload_no_sig(self) -> bytes:
        """messages/proto/helper.go:13-27."""
        return encode_ibft_message(self, with_signature=False)


# ----------------------------------------------------------------------------- encoder
def _varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _f_varint(num: int, v: int) -> bytes:
    return b"" if v == 0 else _varint(num << 3) + _varint(v)


def _f_bytes(num: int, b: bytes) -> bytes:
    return b"" if len(b) == 0 else _varint((num << 3) | 2) + _varint(len(b)) + bytes(b)


def _f_msg(num: int, enc: Optional[bytes]) -> bytes:
    return b"" if enc is None else _varint((num << 3) | 2) + _varint(len(enc)) + enc


def encode_view(v: Optional[View]) -> Optional[bytes]:
    if v is None:
        return None
    return _f_varint(1, v.height) + _f_varint(2, v.round)


def encode_proposal(p: Optional[Proposal]) -> Optional[bytes]:
    if p is None:
        return None
    return _f_bytes(1, p.raw_proposal) + _f_varint(2, p.round)


def encode_pc(pc: Optional[PreparedCertificate]) -> Optional[bytes]:
    if pc is None:
        return None
    out = _f_msg(1, None if pc.proposal_message is None else encode_ibft_message(pc.proposal_message))
    for m in pc.prepare_messages or []:
        out += _f_msg(2, encode_ibft_message(m))
    return out


def encode_rcc(rcc: Optional[RoundChangeCertificate]) -> Optional[bytes]:
    if rcc is None:
        return None
    out = b""
    for m in rcc.round_change_messages:
        out += _f_msg(1, encode_ibft_message(m))
    return out


def encode_payload(p: Payload) -> bytes:
    if p is None:
        return b""
    if isinstance(p, PrePrepareMessage):
        body = _f_msg(1, encode_proposal(p.proposal)) + _f_bytes(2, p.proposal_hash) + _f_msg(3, encode_rcc(p.certificate))
        return _f_msg(5, body)
    if isinstance(p, PrepareMessage):
        return _f_msg(6, _f_bytes(1, p.proposal_hash))
    if isinstance(p, CommitMessage):
        return _f_msg(7, _f_bytes(1, p.proposal_hash) + _f_bytes(2, p.committed_seal))
    if isinstance(p, RoundChangeMessage):
        body = _f_msg(1, encode_proposal(p.last_prepared_proposal)) + _f_msg(2, encode_pc(p.latest_prepared_certificate))
        return _f_msg(8, body)
    raise TypeError(type(p))


def encode_ibft_message(m: IbftMessage, with_signature: bool = True) -> bytes:
    return (
        _f_msg(1, encode_view(m.view))
        + _f_bytes(2, m.from_)
        + (_f_bytes(3, m.signature) if with_signature else b"")
        + _f_varint(4, m.type)
        + encode_payload(m.payload)
    )


# ----------------------------------------------------------------------------- decoder
class DecodeError(ValueError):
    pass


def _read_varint(buf: bytes, pos: int):
    shift = 0
    val = 0
    while True:
        if pos >= len(buf) or shift > 63:
            raise DecodeError("truncated/overlong varint")
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            return val & ((1 << 64) - 1), pos


def _fields(buf: bytes):
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if num == 0:
            raise DecodeError("field number 0")
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            if pos + ln > len(buf):
                raise DecodeError("truncated bytes")
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 1:
            if pos + 8 > len(buf):
                raise DecodeError("truncated fixed64")
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 5:
            if pos + 4 > len(buf):
                raise DecodeError("truncated fixed32")
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise DecodeError(f"unsupported wire type {wt}")
        yield num, wt, v


def decode_view(buf: bytes) -> View:
    v = View()
    for num, wt, val in _fields(buf):
        if num == 1 and wt == 0:
            v.height = val
        elif num == 2 and wt == 0:
            v.round = val
    return v


def decode_proposal(buf: bytes) -> Proposal:
    p = Proposal()
    for num, wt, val in _fields(buf):
        if num == 1 and wt == 2:
            p.raw_proposal = bytes(val)
        elif num == 2 and wt == 0:
            p.round = val
    return p


def decode_pc(buf: bytes) -> PreparedCertificate:
    pc = PreparedCertificate()
    for num, wt, val in _fields(buf):
        if num == 1 and wt == 2:
            pc.proposal_message = decode_ibft_message(val)
        elif num == 2 and wt == 2:
            if pc.prepare_messages is None:
                pc.prepare_messages = []
            pc.prepare_messages.append(decode_ibft_message(val))
    return pc


def decode_rcc(buf: bytes) -> RoundChangeCertificate:
    rcc = RoundChangeCertificate()
    for num, wt, val in _fields(buf):
        if num == 1 and wt == 2:
            rcc.round_change_messages.append(decode_ibft_message(val))
    return rcc


def decode_ibft_message(buf: bytes) -> IbftMessage:
    m = IbftMessage()
    for num, wt, val in _fields(bytes(buf)):
        if num == 1 and wt == 2:
            m.view = decode_view(val)
        elif num == 2 and wt == 2:
            m.from_ = bytes(val)
        elif num == 3 and wt == 2:
            m.signature = bytes(val)
        elif num == 4 and wt == 0:
            m.type = val
        elif num == 5 and wt == 2:
            pp = PrePrepareMessage()
            for n2, w2, v2 in _fields(val):
                if n2 == 1 and w2 == 2:
                    pp.proposal = decode_proposal(v2)
                elif n2 == 2 and w2 == 2:
                    pp.proposal_hash = bytes(v2)
                elif n2 == 3 and w2 == 2:
                    pp.certificate = decode_rcc(v2)
            m.payload = pp
        elif num == 6 and wt == 2:
            pr = PrepareMessage()
            for n2, w2, v2 in _fields(val):
                if n2 == 1 and w2 == 2:
                    pr.proposal_hash = bytes(v2)
            m.payload = pr
        elif num == 7 and wt == 2:
            cm = CommitMessage()
            for n2, w2, v2 in _fields(val):
                if n2 == 1 and w2 == 2:
                    cm.proposal_hash = bytes(v2)
                elif n2 == 2 and w2 == 2:
                    cm.committed_seal = bytes(v2)
            m.payload = cm
        elif num == 8 and wt == 2:
            rc = RoundChangeMessage()
            for n2, w2, v2 in _fields(val):
                if n2 == 1 and w2 == 2:
                    rc.last_prepared_proposal = decode_proposal(v2)
                elif n2 == 2 and w2 == 2:
                    rc.latest_prepared_certificate = decode_pc(v2)
            m.payload = rc
    return m
