"""Generate the committed workload fixtures (tests/golden/*.npz) with the oracle.

    python tests/golden/make_workloads.py
config2.npz  -- SURVEY.md §8(d) config 2 (1k validators, PREPARE + COMMIT, 3,000 signatures, 1 % adversarial)
config3.npz  -- config 3 (10k validators, weighted powers, COMMIT: 10,000 seals + 10,000 sender signatures)
Each holds: items (n x 128 bytes, the packed tuples of include/ibft_verify.h), arena, addrs, powers, group names,
the oracle's verdict bitmap, and the adversarial tags.  bench.py loads config3.npz (it must not import oracle/).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import workloads as wl  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, w):
    bm = wl.oracle_bitmap(w)
    np.savez_compressed(os.path.join(HERE, name),
                        items=w["items"].view(np.uint8).reshape(-1, 128), arena=np.frombuffer(w["arena"], np.uint8),
                        addrs=w["addrs"], powers=w["powers"], groups=np.array(w["groups"]), bitmap=bm,
                        tags=np.array(w["tags"]), proposal_hash=np.frombuffer(w["proposal_hash"], np.uint8),
                        raw_proposal=np.frombuffer(w["raw_proposal"], np.uint8),
                        meta=np.array([w["seed"], w["n"], w["height"], w["round"]], dtype=np.int64))
    n = len(w["items"])
    ok = int(sum(bin(int(x)).count("1") for x in bm))
    print(name, "items", n, "valid", ok, "invalid", n - ok)


if __name__ == "__main__":
    t = time.time()
    save("config2.npz", wl.config2())
    print("config2 %.1fs" % (time.time() - t))
    t = time.time()
    save("config3.npz", wl.config3())
    print("config3 %.1fs" % (time.time() - t))
