"""GPU batched signer (MessageConstructor side): byte-identical to the oracle's signer for the same
nonce, and signatures made with the derived-nonce mode verify through the GPU recover path and the oracle."""
import random

import numpy as np
import pytest

import ibft_b200 as ib
import workloads as wl
from oracle import coracle as co
from oracle import secp256k1 as ec

pytestmark = pytest.mark.gpu


def test_sign_with_given_nonce_matches_oracle(engine):
    rnd = random.Random(31)
    keys = [rnd.getrandbits(256) % (ec.N - 1) + 1 for _ in range(500)] + [1, ec.N - 1]
    digs = [co.keccak256(i.to_bytes(4, "big")) for i in range(len(keys))]
    ks = [ec.rfc6979_k(d, z) for d, z in zip(keys, digs)]
    ks[0], ks[1] = ec.N - 1, 1
    got = engine.sign_batch(keys, digs, ks)
    for d, z, k, g in zip(keys, digs, ks, got):
        assert g == co.sign_with_k(d, z, k, True)
    # unusable nonces -> all-zero signature, never a bogus one
    bad = engine.sign_batch([5, 5, 5], [digs[0]] * 3, [0, ec.N, ec.N + 7])
    assert bad == [bytes(65)] * 3
    assert engine.sign_batch([], []) == []


def test_derived_nonce_signatures_verify_on_gpu_and_oracle(engine):
    vs = wl.ValidatorSet(41, 300)
    engine.set_validators(6, 3, vs.addr_array(), None)
    ph = co.keccak256(b"block")
    sd = wl.seal_digest(ph)
    seals = engine.sign_batch(vs.keys, [sd] * vs.n)
    assert len(set(seals)) == vs.n and all(s != bytes(65) for s in seals)
    assert engine.sign_batch(vs.keys[:10], [sd] * 10) == seals[:10]              # deterministic
    for s in seals:
        assert int.from_bytes(s[32:64], "big") <= ec.N // 2                       # low-s
    items = np.concatenate([wl.make_item(seals[i], vs.addrs[i], 2, ph, 0) for i in range(vs.n)])
    g = engine.groups(1, slot=6)
    bitmap, results, _ = engine.verify_batch(items, b"", g)
    assert sum(bin(int(w)).count("1") for w in bitmap) == vs.n and bool(results[0]["has_quorum"])
    for i in (0, 17, 299):
        assert co.ecrecover_address(sd, seals[i]) == vs.addrs[i]
