"""BASELINE configs 4 and 5 at their stated size, CPU side: the regenerated inputs match the committed fingerprints and the
oracle reproduces the committed answers (tests/golden/config{4_n10k,5}_pin.npz, generator tests/golden/make_pins.py)."""
import numpy as np

import workloads as wl
from oracle import coracle as co
from oracle import ibft_proto as ip


def test_config5_regenerates_to_the_pinned_bytes_and_oracle_bitmap():
    w, pin = wl.load_full("config5")
    assert len(w["items"]) == int(pin["meta"][0]) and w["n_messages"] == 100_000 and len(w["tables"]) == 16
    assert all(len(t) == 10_000 for t in w["tables"])
    bm = co.verify_batch(w["items"], w["arena"], tables=w["tables"], group_table=w["group_table"], n_threads=wl.N_THREADS)
    assert np.array_equal(bm, pin["bitmap"])
    # the 45/45/9/1 mix, and ~1 % of the tuples invalid
    kinds = np.bincount(w["items"]["group"][:100_000] % 4, minlength=4) / 100_000
    assert abs(kinds[ip.PREPARE] - 0.45) < 0.01 and abs(kinds[ip.COMMIT] - 0.45) < 0.01 and abs(kinds[ip.ROUND_CHANGE] - 0.09) < 0.01
    ones = sum(bin(int(x)).count("1") for x in bm)
    assert 0.985 < ones / len(w["items"]) < 0.995


def test_config4_n10k_regenerates_to_the_pinned_bytes_and_decisions():
    w, pin = wl.load_full("config4_n10k")
    assert w["n"] == 10_000 and w["quorum"] == 6_667 and len(w["items"]) == int(pin["meta"][0])
    bm = co.verify_batch(w["items"], w["arena"], tables=[w["addrs"]], group_table=[0], n_threads=wl.N_THREADS)
    assert np.array_equal(bm, pin["bitmap"])
    valid, hq = wl.config4_expected(w, bm)
    assert np.array_equal(np.packbits(valid), pin["rc_valid"]) and int(hq) == int(pin["has_quorum"][0])
    assert int(valid.sum()) == 10_000 - 100 - 1      # 100 corrupted nested signatures + the certificate below quorum


def test_round_change_head_plus_certificate_is_payload_no_sig():
    """the identity IBFT_KIND_PAYLOAD2 rests on, against the codec itself (small case)"""
    w = wl.config4_n10k(n=40, n_bad=3)
    for i in range(40):
        c = int(w["cert_of"][i])
        if c < 0:
            continue
        full = ip.IbftMessage(ip.View(w["height"], 1), bytes(w["addrs"][i]), b"", ip.ROUND_CHANGE,
                              ip.RoundChangeMessage(ip.Proposal(w["raw_proposal"], 0), w["variants"][c])).payload_no_sig()
        assert full == w["heads"][i] + w["blobs"][c]
        it = w["items"][i]
        assert int(it["kind"]) == wl.KIND_PAYLOAD2
        assert co.ecrecover_address(co.keccak256(full), bytes(w["rc_sigs"][i])) == bytes(w["addrs"][i])
