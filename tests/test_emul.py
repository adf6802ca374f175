"""Host emulation of the DEVICE arithmetic (the very csrc/*.cuh headers, portable path) against the oracle.

Runs on a CPU-only box: it validates the algorithms the kernels use -- lazy-reduced field arithmetic, GLV split,
Booth digits, the group law with its exceptional cases, Keccak, address derivation -- before any GPU time is spent.
The PTX carry-chain multipliers themselves are covered on the GPU by test_gpu_primitives.py."""
import ctypes
import os
import random

import numpy as np

from oracle import coracle as co
from oracle import secp256k1 as ec
from oracle.keccak import keccak256

P, N = ec.P, ec.N
LAM = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72
EDGE = [0, 1, 2, P - 1, P, P + 1, 2**256 - 1, 2**255, N, N - 1, 2**32, 2**32 + 977, 2**256 - 2**32 - 978, (1 << 224) - 1]


def B(b):
    return (ctypes.c_uint8 * len(b)).from_buffer_copy(b)


def dbg(E, op, a, b=0, c=None, outn=32):
    out = (ctypes.c_uint8 * 64)()
    rc = E.emul_debug_op(op, B(a.to_bytes(32, "big")), B(b.to_bytes(32, "big")), B(c) if c else None, out)
    return rc, bytes(out)[:outn]


def test_field_and_scalar_ops(emul):
    rnd = random.Random(3)
    vals = EDGE + [rnd.getrandbits(256) for _ in range(150)]
    for a in vals:
        for b in rnd.sample(vals, 4) + EDGE[:7]:
            assert int.from_bytes(dbg(emul, 1, a, b)[1], "big") == a * b % P
            assert int.from_bytes(dbg(emul, 8, a, b)[1], "big") == (a + b) % P
            assert int.from_bytes(dbg(emul, 9, a, b)[1], "big") == (a - b) % P
            assert int.from_bytes(dbg(emul, 5, a, b)[1], "big") == a * b % N
        assert int.from_bytes(dbg(emul, 2, a)[1], "big") == a * a % P
    for a in vals[:40]:
        assert int.from_bytes(dbg(emul, 3, a)[1], "big") == (pow(a % P, -1, P) if a % P else 0)
        assert int.from_bytes(dbg(emul, 6, a)[1], "big") == (pow(a % N, -1, N) if a % N else 0)
        assert int.from_bytes(dbg(emul, 4, a)[1], "big") == pow(a % P, (P + 1) // 4, P)


def test_glv_split(emul):
    rnd = random.Random(4)
    ks = [0, 1, 2, N - 1, N - 2, LAM, N - LAM, (N + 1) // 2, N // 2, 2**128, 2**128 - 1, 2**255] + [rnd.getrandbits(256) % N for _ in range(1500)]
    for k in ks:
        _, o = dbg(emul, 10, k, outn=64)
        k1 = int.from_bytes(o[0:20], "little") * (-1 if o[20] else 1)
        k2 = int.from_bytes(o[24:44], "little") * (-1 if o[44] else 1)
        assert (k1 + k2 * LAM) % N == k and abs(k1) < 2**129 and abs(k2) < 2**129


def ecm(E, a, b, pt):
    rc, o = dbg(E, 7, a, b, pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big"), 64)
    return None if rc else (int.from_bytes(o[:32], "big"), int.from_bytes(o[32:], "big"))


def test_ecmult_including_exceptional_cases(emul):
    rnd = random.Random(5)
    G = ec.G
    pt = ec.point_mul(12345, G)
    cases = [(0, 0), (1, 0), (0, 1), (2, 0), (0, 2), (N - 1, 0), (0, N - 1), (1, N - 1), (5, 7), (N - 1, N - 1), (LAM, 0), (0, LAM), (N - LAM, LAM)]
    cases += [(rnd.getrandbits(256) % N, rnd.getrandbits(256) % N) for _ in range(12)]
    for a, b in cases:
        assert ecm(emul, a, b, pt) == ec.point_add(ec.point_mul(a, G), ec.point_mul(b, pt))
    # P chosen so that partial sums collide with table entries: doubling / cancellation branches of the adders
    for p2 in [G, ec.point_neg(G), ec.point_mul(2, G), ec.point_mul(LAM, G), ec.point_mul(N - LAM, G), ec.point_mul(8, G), ec.point_mul(N - 8, G)]:
        for a, b in [(1, 1), (1, N - 1), (2, N - 1), (N - 2, 1), (8, 1), (8, N - 1), (3, 5), (LAM, 1), (1, LAM), (N - 1, N - 1), (7, 1), (16, N - 2)]:
            assert ecm(emul, a, b, p2) == ec.point_add(ec.point_mul(a, G), ec.point_mul(b, p2)), (a, b)


def emul_item(E, item, arena=b""):
    out = (ctypes.c_uint8 * 20)()
    a = np.frombuffer(arena, np.uint8) if arena else np.zeros(1, np.uint8)
    ok = E.emul_verify_item(item.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(arena)), out)
    return bool(ok), bytes(out)


def test_recover_valid_highs_and_corrupted(emul):
    import workloads as wl
    rnd = random.Random(6)
    for i in range(60):
        d = rnd.getrandbits(256) % (N - 1) + 1
        dig = keccak256(bytes([i]) * 3)
        sig = wl.sign(d, dig, low_s=bool(i & 1))
        if i % 3 == 0:
            s = int.from_bytes(sig[32:64], "big")
            sig = sig[:32] + (N - s).to_bytes(32, "big") + bytes([sig[64] ^ 1])
        addr = co.ecrecover_address(dig, sig)
        assert emul_item(emul, wl.make_item(sig, addr, 0, dig)) == (True, addr)
        bad = bytearray(sig)
        bad[rnd.randrange(64)] ^= 1 << rnd.randrange(8)
        exp = co.ecrecover_address(dig, bytes(bad))
        assert emul_item(emul, wl.make_item(bytes(bad), addr, 0, dig)) == (exp == addr, exp or bytes(20))
        ph = keccak256(bytes([i, 1]))
        sg = wl.sign(d, wl.seal_digest(ph))
        assert emul_item(emul, wl.make_item(sg, addr, 2, ph)) == (True, addr)


def test_config2_fixture_bit_exact(emul):
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "config2.npz"))
    items = np.ascontiguousarray(d["items"]).view(co.ITEM_DTYPE).reshape(-1)
    arena = d["arena"].tobytes()
    members = {bytes(a) for a in d["addrs"]}
    bm = d["bitmap"]
    for i in range(len(items)):
        ok, _ = emul_item(emul, items[i:i + 1], arena)
        ok = ok and bytes(items[i]["signer"]) in members
        assert ok == bool((bm[i >> 5] >> (i & 31)) & 1), (i, d["tags"][i])
