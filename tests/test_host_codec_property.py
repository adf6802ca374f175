"""Property test (hypothesis): random IbftMessage trees -- the C++ codec (host/proto.hpp) and the oracle codec
(oracle/ibft_proto.py, itself pinned to the reference descriptor) must produce identical wire bytes and PayloadNoSig."""
import importlib

from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import ibft_proto as ip

host = importlib.import_module("go-ibft_b200.host")

u64 = st.integers(min_value=0, max_value=2**64 - 1)
small_bytes = st.binary(max_size=40)
views = st.one_of(st.none(), st.builds(ip.View, u64, u64))
proposals = st.one_of(st.none(), st.builds(ip.Proposal, st.binary(max_size=200), u64))


def messages(depth):
    leaf_payloads = [st.none(), st.builds(ip.PrepareMessage, small_bytes), st.builds(ip.CommitMessage, small_bytes, st.binary(max_size=70))]
    if depth <= 0:
        payload = st.one_of(*leaf_payloads, st.builds(ip.PrePrepareMessage, proposals, small_bytes, st.none()),
                            st.builds(ip.RoundChangeMessage, proposals, st.none()))
    else:
        inner = messages(depth - 1)
        pcs = st.one_of(st.none(), st.builds(ip.PreparedCertificate, st.one_of(st.none(), inner),
                                             st.one_of(st.none(), st.lists(inner, min_size=1, max_size=3))))
        rccs = st.one_of(st.none(), st.builds(ip.RoundChangeCertificate, st.lists(inner, max_size=3)))
        payload = st.one_of(*leaf_payloads, st.builds(ip.PrePrepareMessage, proposals, small_bytes, rccs),
                            st.builds(ip.RoundChangeMessage, proposals, pcs))
    return st.builds(ip.IbftMessage, views, small_bytes, st.binary(max_size=70), st.integers(min_value=0, max_value=3), payload)


@settings(max_examples=300, deadline=None)
@given(messages(2))
def test_cpp_codec_equals_oracle_codec(m):
    wire = ip.encode_ibft_message(m)
    assert host.reencode(wire, True) == wire
    assert host.reencode(wire, False) == m.payload_no_sig()
    # and the oracle decoder round-trips its own encoding
    assert ip.encode_ibft_message(ip.decode_ibft_message(wire)) == wire


@settings(max_examples=200, deadline=None)
@given(st.binary(max_size=120))
def test_cpp_decoder_never_crashes_on_garbage(blob):
    out = host.reencode(blob, True)
    try:
        want = ip.encode_ibft_message(ip.decode_ibft_message(blob))
    except ip.DecodeError:
        want = None
    except RecursionError:
        return
    assert out == want
