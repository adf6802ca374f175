"""Pins of the full-size BASELINE configs 4 and 5 (tests/workloads.py config4_n10k / config5_full).

    python tests/golden/make_pins.py
The inputs are regenerated deterministically wherever the tests run (seconds, oracle bulk generators); what is COMMITTED is
their fingerprint -- SHA-256 of every input array -- and the oracle's answers for them: the verdict bitmap, and for config 5 the
per-group quorum results (n_valid, n_distinct, 320-bit power, has_quorum), for config 4 the per-ROUND_CHANGE validity and the
round's quorum decision.  A test that regenerates different bytes fails on the fingerprint before it compares any verdict.
"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import workloads as wl  # noqa: E402
from oracle import coracle as co  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a) -> np.ndarray:
    b = a if isinstance(a, (bytes, bytearray)) else np.ascontiguousarray(a).tobytes()
    return np.frombuffer(hashlib.sha256(b).digest(), np.uint8)


def group_results(w, bitmap):
    """HasQuorum restated with Python big ints (core/validator_manager.go:77-96, :130-135) per group of config 5."""
    bits = np.unpackbits(bitmap.view(np.uint8), bitorder="little")[: len(w["items"])]
    pw = [int.from_bytes(bytes(p), "big") for p in w["powers"]]
    quorum = 2 * sum(pw) // 3 + 1
    out = np.zeros((w["n_groups"], 4 + 5), dtype=np.uint64)  # n_valid, n_distinct, has_quorum, (pad), power limbs
    idx = [{bytes(a): i for i, a in enumerate(t)} for t in w["tables"]]
    items = w["items"]
    for g in range(w["n_groups"]):
        sel = np.nonzero((items["group"] == g) & (bits == 1))[0]
        voters = {idx[w["group_table"][g]][bytes(items["signer"][i])] for i in sel}
        power = sum(pw[v] for v in voters)
        out[g, 0], out[g, 1], out[g, 2] = len(sel), len(voters), int(power >= quorum)
        for j in range(5):
            out[g, 4 + j] = (power >> (64 * j)) & ((1 << 64) - 1)
    return out


def pin_config5():
    t = time.time()
    w = wl.config5_full()
    t_gen = time.time() - t
    t = time.time()
    bm = co.verify_batch(w["items"], w["arena"], tables=w["tables"], group_table=w["group_table"], n_threads=wl.N_THREADS)
    t_ver = time.time() - t
    np.savez_compressed(os.path.join(HERE, "config5_pin.npz"), sha_items=sha(w["items"]), sha_arena=sha(w["arena"]),
                        sha_tables=sha(np.concatenate(w["tables"])), bitmap=bm, results=group_results(w, bm),
                        meta=np.array([len(w["items"]), len(w["arena"]), w["n_messages"], w["n_groups"]], dtype=np.int64))
    ok = int(sum(bin(int(x)).count("1") for x in bm))
    print("config5: %d tuples, %d valid, arena %d B; generate %.1fs, oracle verify %.1fs" % (len(w["items"]), ok, len(w["arena"]), t_gen, t_ver))


def pin_config4():
    t = time.time()
    w = wl.config4_n10k()
    t_gen = time.time() - t
    t = time.time()
    bm = co.verify_batch(w["items"], w["arena"], tables=[w["addrs"]], group_table=[0], n_threads=wl.N_THREADS)
    t_ver = time.time() - t
    valid, hq = wl.config4_expected(w, bm)
    np.savez_compressed(os.path.join(HERE, "config4_n10k_pin.npz"), sha_items=sha(w["items"]), sha_arena=sha(w["arena"]),
                        sha_table=sha(w["addrs"]), bitmap=bm, rc_valid=np.packbits(valid), has_quorum=np.array([int(hq)]),
                        meta=np.array([len(w["items"]), len(w["arena"]), w["n"], w["quorum"], int(valid.sum())], dtype=np.int64))
    ok = int(sum(bin(int(x)).count("1") for x in bm))
    print("config4: %d tuples, %d valid, arena %d B, %d of %d ROUND_CHANGE valid, quorum %s; generate %.1fs, oracle verify %.1fs"
          % (len(w["items"]), ok, len(w["arena"]), int(valid.sum()), w["n"], hq, t_gen, t_ver))


if __name__ == "__main__":
    pin_config5()
    pin_config4()
