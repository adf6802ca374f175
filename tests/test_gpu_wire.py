"""Raw gossip frames on the device (IBFT_KIND_WIRE / IBFT_KIND_WIRE_SEAL): the kernel parses the proto3
frame, cuts the signature TLV out to obtain PayloadNoSig, and verifies -- bit-exact against the oracle codec + verifier,
including non-canonical and mutated frames (which must come back as NEEDS_HOST, never as a guessed verdict)."""
import random

import numpy as np
import pytest

import ibft_b200 as ib
from test_emul import sample_frames, wire_expectation, wire_item

pytestmark = pytest.mark.gpu


def test_raw_frames_batch(engine):
    vs, frames = sample_frames()
    members = set(vs.addrs[:5])                       # validator 5 signs correctly but is NOT in the table
    engine.set_validators(7, 1, np.frombuffer(b"".join(vs.addrs[:5]), np.uint8).reshape(5, 20), None)
    rnd = random.Random(9)
    for f in list(frames[:20]):
        for _ in range(6):
            b = bytearray(f)
            i = rnd.randrange(len(b))
            b[i] ^= 1 << rnd.randrange(8)
            frames.append(bytes(b))
    arena = bytearray()
    items, want = [], []
    for wire in frames:
        off = len(arena)
        arena.extend(wire)
        for kind in (ib.KIND_WIRE, ib.KIND_WIRE_SEAL):
            items.append(wire_item(kind, off, len(wire), 0 if kind == ib.KIND_WIRE else 1))
            want.append(wire_expectation(wire, kind, members))
    items = np.concatenate(items).view(ib.ITEM_DTYPE)
    groups = engine.groups(2, slot=7)
    bitmap, results, _ = engine.verify_batch(items, bytes(arena) or b"\x00", groups)
    status = engine.last_item_status(len(items))
    bits = np.unpackbits(bitmap.view(np.uint8), bitorder="little")[: len(items)]
    for i, (st, ok) in enumerate(want):
        assert (int(status[i]), bool(bits[i])) == (st, ok), (i, frames[i // 2].hex())
    # quorum kernels resolve the signer of a raw frame from the frame itself
    for g in (0, 1):
        voters = set()
        for i, (st, ok) in enumerate(want):
            if ok and int(items[i]["group"]) == g:
                from oracle import ibft_proto as ip
                voters.add(ip.decode_ibft_message(frames[i // 2]).from_)
        assert int(results[g]["n_distinct"]) == len(voters) and int(results[g]["n_valid"]) == sum(1 for i, (st, ok) in enumerate(want) if ok and int(items[i]["group"]) == g)
        assert bool(results[g]["has_quorum"]) == (len(voters) >= 2 * 5 // 3 + 1)
    assert sum(1 for st, _ in want if st == 1) > 20 and sum(1 for _, ok in want if ok) > 30
