"""Deterministic synthetic workloads of SURVEY.md §8(d) (configs 2-5), built with the ORACLE's signer.

Test infrastructure: imports oracle/.  The committed fixtures under tests/golden/*.npz are produced from here by
tests/golden/make_workloads.py; bench.py only loads the fixtures (it never imports this module).
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from oracle import coracle as co  # noqa: E402
from oracle import ibft_proto as ip  # noqa: E402
from oracle import secp256k1 as ec  # noqa: E402

ITEM = co.ITEM_DTYPE
KIND_DIGEST, KIND_PAYLOAD, KIND_SEAL, KIND_INVALID = 0, 1, 2, 255


def privkey(seed: int, i: int) -> int:
    """privkey_i = Keccak-256("ibft-b200-validator" || u32_be(seed) || u32_be(i)) mod (n-1) + 1   (SURVEY.md §8d)."""
    h = co.keccak256(b"ibft-b200-validator" + seed.to_bytes(4, "big") + i.to_bytes(4, "big"))
    return int.from_bytes(h, "big") % (ec.N - 1) + 1


def address_of(d: int) -> bytes:
    pub = co.pubkey_from_scalar(d)
    return co.keccak256(pub)[12:]


def sign(d: int, digest: bytes, low_s: bool = True) -> bytes:
    return co.sign_with_k(d, digest, ec.rfc6979_k(d, digest), low_s)


def proposal_hash(raw: bytes, rnd: int) -> bytes:
    """Synthetic convention of SURVEY.md §8(c): Keccak-256(Keccak-256(rawProposal) || u64_be(round))."""
    return co.keccak256(co.keccak256(raw) + rnd.to_bytes(8, "big"))


def seal_digest(phash: bytes) -> bytes:
    return co.keccak256(phash + b"\x02")


class ValidatorSet:
    def __init__(self, seed: int, n: int, weighted: bool = False):
        self.seed, self.n = seed, n
        self.keys = [privkey(seed, i) for i in range(n)]
        self.addrs = [address_of(d) for d in self.keys]
        self.powers = [1 + (i % 7) if weighted else 1 for i in range(n)]

    def addr_array(self) -> np.ndarray:
        return np.frombuffer(b"".join(self.addrs), dtype=np.uint8).reshape(self.n, 20).copy()

    def power_array(self) -> np.ndarray:
        return np.frombuffer(b"".join(p.to_bytes(32, "big") for p in self.powers), dtype=np.uint8).reshape(self.n, 32).copy()


def make_item(sig: bytes, signer: bytes, kind: int, digest: bytes = b"", group: int = 0, off: int = 0, ln: int = 0) -> np.ndarray:
    it = np.zeros(1, dtype=ITEM)
    if len(sig) != 65 or len(signer) != 20:
        it["kind"][0] = KIND_INVALID
        it["group"][0] = group
        return it
    it["r"][0] = np.frombuffer(sig[:32], np.uint8)
    it["s"][0] = np.frombuffer(sig[32:64], np.uint8)
    it["v"][0] = sig[64]
    if digest:
        it["digest"][0] = np.frombuffer(digest, np.uint8)
    it["signer"][0] = np.frombuffer(signer, np.uint8)
    it["kind"][0] = kind
    it["group"][0] = group
    it["payload_off"][0] = off
    it["payload_len"][0] = ln
    return it


ADVERSARIAL_KINDS = ("bad_v", "r_zero", "s_ge_n", "flipped_digest", "non_member", "from_ne_signer")


def corrupt(kind: str, sig: bytes, signer: bytes, other_addr: bytes):
    """Adversarial variants that only touch the signature / expected signer."""
    if kind == "bad_v":
        return sig[:64] + bytes([sig[64] + 2]), signer
    if kind == "r_zero":
        return bytes(32) + sig[32:], signer
    if kind == "s_ge_n":
        s = int.from_bytes(sig[32:64], "big") + ec.N
        return sig[:32] + (s % (1 << 256)).to_bytes(32, "big") + sig[64:], signer
    if kind == "from_ne_signer":
        return sig, other_addr
    raise ValueError(kind)


def build_round(seed: int, n: int, height: int, rnd: int, *, with_prepare: bool, with_commit_sender: bool, with_seals: bool,
                weighted: bool = False, adversarial: bool = True, raw_seed: int = 1):
    """One (height, round) of PREPARE / COMMIT traffic from all n validators.

    1 % adversarial items at indices i % 100 == 7, cycling through ADVERSARIAL_KINDS."""
    vs = ValidatorSet(seed, n, weighted)
    rng = np.random.default_rng(raw_seed)
    raw = rng.integers(0, 256, 1024, dtype=np.uint8).tobytes()
    ph = proposal_hash(raw, rnd)
    sd = seal_digest(ph)
    outsider = privkey(seed + 1000, 0)
    outsider_addr = address_of(outsider)
    items, arena, wire, tags = [], bytearray(), [], []
    groups = []
    view = ip.View(height, rnd)

    def adv_kind(i, salt):
        if not adversarial or i % 100 != 7:
            return None
        return ADVERSARIAL_KINDS[((i // 100) + salt) % len(ADVERSARIAL_KINDS)]

    def add_payload_item(msg: ip.IbftMessage, i: int, group: int, tag):
        key, signer = vs.keys[i], vs.addrs[i]
        if tag == "non_member":
            key, signer = outsider, outsider_addr
            msg = ip.IbftMessage(msg.view, outsider_addr, b"", msg.type, msg.payload)
        payload = msg.payload_no_sig()
        sig = sign(key, co.keccak256(payload))
        if tag == "flipped_digest":
            b = bytearray(payload)
            b[-5] ^= 0x10  # the signed bytes change after signing
            payload = bytes(b)
        elif tag in ("bad_v", "r_zero", "s_ge_n", "from_ne_signer"):
            sig, signer = corrupt(tag, sig, signer, vs.addrs[(i + 1) % n])
        off = len(arena)
        arena.extend(payload)
        items.append(make_item(sig, signer, KIND_PAYLOAD, b"", group, off, len(payload)))
        tags.append(tag or "")
        return sig

    if with_prepare:
        g = len(groups)
        groups.append("PREPARE")
        for i in range(n):
            m = ip.IbftMessage(view, vs.addrs[i], b"", ip.PREPARE, ip.PrepareMessage(ph))
            sig = add_payload_item(m, i, g, adv_kind(i, 0))
            wire.append(ip.encode_ibft_message(ip.IbftMessage(view, vs.addrs[i], sig, ip.PREPARE, ip.PrepareMessage(ph))))
    if with_commit_sender or with_seals:
        gs = len(groups)
        if with_commit_sender:
            groups.append("COMMIT")
        gseal = len(groups)
        if with_seals:
            groups.append("COMMIT_SEAL")
        for i in range(n):
            seal = sign(vs.keys[i], sd)
            if with_seals:
                tag = adv_kind(i, 1)
                seal_signer, seal_hash, seal_sig = vs.addrs[i], ph, seal
                if tag == "non_member":
                    seal_sig, seal_signer = sign(outsider, sd), outsider_addr
                elif tag == "flipped_digest":
                    b = bytearray(ph)
                    b[7] ^= 0x10
                    seal_hash = bytes(b)
                elif tag is not None:
                    seal_sig, seal_signer = corrupt(tag, seal, seal_signer, vs.addrs[(i + 1) % n])
                items.append(make_item(seal_sig, seal_signer, KIND_SEAL, seal_hash, gseal))
                tags.append(tag or "")
            if with_commit_sender:
                m = ip.IbftMessage(view, vs.addrs[i], b"", ip.COMMIT, ip.CommitMessage(ph, seal))
                sig = add_payload_item(m, i, gs, adv_kind(i, 2))
                wire.append(ip.encode_ibft_message(ip.IbftMessage(view, vs.addrs[i], sig, ip.COMMIT, ip.CommitMessage(ph, seal))))
    return dict(items=np.concatenate(items), arena=bytes(arena), addrs=vs.addr_array(), powers=vs.power_array(), groups=groups,
                wire=wire, tags=tags, raw_proposal=raw, proposal_hash=ph, height=height, round=rnd, seed=seed, n=n)


def oracle_bitmap(w, n_threads: int = 8) -> np.ndarray:
    gt = [0] * len(w["groups"])
    return co.verify_batch(w["items"], w["arena"], tables=[w["addrs"]], group_table=gt, n_threads=n_threads)


def config2():
    """1k-validator PREPARE+COMMIT batch: 3,000 signatures (1k PREPARE sender, 1k COMMIT sender, 1k seals)."""
    return build_round(1, 1000, 1_000_000, 0, with_prepare=True, with_commit_sender=True, with_seals=True, raw_seed=1)


def config3(weighted: bool = True):
    """10k-validator COMMIT round: 10,000 seals + 10,000 sender signatures."""
    return build_round(2, 10_000, 1_000_000, 0, with_prepare=False, with_commit_sender=True, with_seals=True,
                       weighted=weighted, raw_seed=2)


# =====================================================================================================================
# BASELINE.json configs 4 and 5 at their STATED size (10,000-validator tables).  Built with the oracle's bulk generators
# (oracle/coracle.py: privkeys / addresses / sign_derived_batch / keccak256_batch -- the deterministic Keccak-derived nonce
# k = Keccak-256(d || z || ctr), so the bytes are reproducible without RFC 6979's HMAC) in seconds.  The inputs are too large to
# commit (config 5 is 29 MB of incompressible signatures): they are REGENERATED at test time and pinned by the SHA-256 of every
# array plus the oracle's verdict bitmap, committed in tests/golden/config{4_n10k,5}_pin.npz (tests/golden/make_pins.py).
# =====================================================================================================================
KIND_PAYLOAD2 = 5
N_THREADS = min(32, os.cpu_count() or 8)


def _fill_items(n):
    return np.zeros(n, dtype=ITEM)


def _put_sigs(items, sigs, signers):
    items["r"] = sigs[:, :32]
    items["s"] = sigs[:, 32:64]
    items["v"] = sigs[:, 64]
    items["signer"] = signers


def _corrupt_in_place(items, idx, kind, other_signer):
    """adversarial variants of SURVEY.md §8d config 2 applied to an already signed tuple"""
    if kind == "bad_v":
        items["v"][idx] += 2
    elif kind == "r_zero":
        items["r"][idx] = 0
    elif kind == "s_ge_n":
        s = int.from_bytes(bytes(items["s"][idx]), "big") + ec.N
        items["s"][idx] = np.frombuffer((s % (1 << 256)).to_bytes(32, "big"), np.uint8)
    elif kind == "from_ne_signer":
        items["signer"][idx] = other_signer
    else:
        raise ValueError(kind)


def config5_full(n_val: int = 10_000, total: int = 100_000, n_heights: int = 16):
    """BASELINE config 5: `total` pending messages across `n_heights` concurrent heights (1,000,000 ...), per-height validator
    tables of `n_val` (seeds 10 ..., weighted powers 1 + i mod 7), mix 45 % PREPARE / 45 % COMMIT (+ committed seal) / 9 %
    ROUND_CHANGE (no certificate) / 1 % PREPREPARE.  One tuple per sender signature, in arrival order, then one
    per committed seal (~145k tuples).  Groups: g = 4*k + type for the sender signatures of height k, 4*n_heights + k for its
    seals.  1 % of the messages (index % 100 == 7) are adversarial, cycling through the six kinds of config 2; index % 1000 ==
    13 is a message replayed under another height's view (the signed bytes differ -> invalid)."""
    heights = [1_000_000 + k for k in range(n_heights)]
    rng = np.random.default_rng(5)
    privs = [co.privkeys(10 + k, n_val) for k in range(n_heights)]
    addrs = [co.addresses(p, N_THREADS) for p in privs]
    powers = np.frombuffer(b"".join((1 + (i % 7)).to_bytes(32, "big") for i in range(n_val)), np.uint8).reshape(n_val, 32).copy()
    raw = rng.integers(0, 256, 1024, dtype=np.uint8).tobytes()
    ph = proposal_hash(raw, 0)
    sd = seal_digest(ph)
    types = rng.choice([ip.PREPARE, ip.COMMIT, ip.ROUND_CHANGE, ip.PREPREPARE], size=total, p=[0.45, 0.45, 0.09, 0.01])
    hk = rng.integers(0, n_heights, size=total)
    vi = rng.integers(0, n_val, size=total)
    outsider_priv = co.privkeys(9999, 1)
    outsider_addr = bytes(co.addresses(outsider_priv, 1)[0])

    def tag_of(i):
        if i % 1000 == 13:
            return "replayed"
        return ADVERSARIAL_KINDS[(i // 100) % len(ADVERSARIAL_KINDS)] if i % 100 == 7 else ""
    tags = [tag_of(i) for i in range(total)]
    # committed seals first: the COMMIT payload embeds them
    ci = np.nonzero(types == ip.COMMIT)[0]
    seal_privs = np.stack([outsider_priv[0] if tags[i] == "non_member" else privs[hk[i]][vi[i]] for i in ci])
    seal_sigs = co.sign_derived_batch(seal_privs, np.tile(np.frombuffer(sd, np.uint8), (len(ci), 1)), N_THREADS)
    seal_of = {int(i): bytes(seal_sigs[j]) for j, i in enumerate(ci)}
    # PayloadNoSig of every message: `shown` is what the verifier gets, `signed_digest` what the sender signed
    arena = bytearray()
    offs, lens = np.zeros(total, np.uint32), np.zeros(total, np.uint32)
    signed_digest = np.zeros((total, 32), np.uint8)
    sign_priv = np.zeros((total, 32), np.uint8)
    signer = np.zeros((total, 20), np.uint8)
    for i in range(total):
        k, v, t, tag = int(hk[i]), int(vi[i]), int(types[i]), tags[i]
        frm = outsider_addr if tag == "non_member" else bytes(addrs[k][v])
        if t == ip.PREPARE:
            body = ip.PrepareMessage(ph)
        elif t == ip.COMMIT:
            body = ip.CommitMessage(ph, seal_of[i])
        elif t == ip.ROUND_CHANGE:
            body = ip.RoundChangeMessage(None, None)
        else:
            body = ip.PrePrepareMessage(ip.Proposal(raw, 0), ph, None)
        p = ip.IbftMessage(ip.View(heights[k], 0), frm, b"", t, body).payload_no_sig()
        signed_digest[i] = np.frombuffer(co.keccak256(p), np.uint8)
        if tag == "replayed":
            p = ip.IbftMessage(ip.View(heights[(k + 1) % n_heights], 0), frm, b"", t, body).payload_no_sig()
        elif tag == "flipped_digest":
            b = bytearray(p)
            b[-5] ^= 0x10
            p = bytes(b)
        offs[i], lens[i] = len(arena), len(p)
        arena.extend(p)
        sign_priv[i] = outsider_priv[0] if tag == "non_member" else privs[k][v]
        signer[i] = np.frombuffer(frm, np.uint8)
    sigs = co.sign_derived_batch(sign_priv, signed_digest, N_THREADS)
    items = _fill_items(total + len(ci))
    _put_sigs(items[:total], sigs, signer)
    items["kind"][:total] = KIND_PAYLOAD
    items["group"][:total] = 4 * hk + types
    items["payload_off"][:total] = offs
    items["payload_len"][:total] = lens
    for i in range(total):
        if tags[i] in ("bad_v", "r_zero", "s_ge_n", "from_ne_signer"):
            _corrupt_in_place(items, i, tags[i], addrs[int(hk[i])][(int(vi[i]) + 1) % n_val])
    # committed seals (IsValidCommittedSeal at the COMMIT's height): tuple total + j belongs to COMMIT message ci[j]
    seal_signer = np.stack([np.frombuffer(outsider_addr, np.uint8) if tags[i] == "non_member" else addrs[hk[i]][vi[i]] for i in ci])
    sl = items[total:]
    _put_sigs(sl, seal_sigs, seal_signer)
    sl["kind"] = KIND_SEAL
    sl["digest"] = np.frombuffer(ph, np.uint8)
    sl["group"] = 4 * n_heights + hk[ci]
    seal_tags = []
    for j, i in enumerate(ci):
        tag = tags[i] if tags[i] != "replayed" else ""   # a seal carries no view
        if tag == "flipped_digest":
            sl["digest"][j][7] ^= 0x10
        elif tag in ("bad_v", "r_zero", "s_ge_n", "from_ne_signer"):
            _corrupt_in_place(sl, j, tag, addrs[int(hk[i])][(int(vi[i]) + 1) % n_val])
        seal_tags.append(tag)
    n_groups = 5 * n_heights
    group_table = [g // 4 for g in range(4 * n_heights)] + list(range(n_heights))
    return dict(items=items, arena=bytes(arena), tables=addrs, powers=powers, heights=heights, n_groups=n_groups,
                group_table=group_table, tags=tags + seal_tags, n_messages=total, proposal_hash=ph, raw_proposal=raw)


def config4_n10k(n: int = 10_000, n_bad: int = 100):
    """BASELINE config 4 at its stated size, dedup mode: n validators (seed 3, unit power, quorum 6,667 at
    n = 10,000), height 1,000,000, round 1: every validator sends a ROUND_CHANGE that embeds one of three prepared certificates
    of round 0 (the same PREPREPARE of validator 0 + a window of Q-1 PREPAREs of the others).  `n_bad` of the messages embed a
    certificate with ONE corrupted nested PREPARE signature instead (three corrupted variants, one per certificate); message 11
    embeds a certificate below quorum and message 5 no certificate at all.
    The unique signatures -- n ROUND_CHANGE senders + 1 PREPREPARE + the distinct PREPAREs + the corrupted variants -- are the
    tuples.  A ROUND_CHANGE's signed bytes end with its whole certificate (909 KB at n = 10,000): its tuple is IBFT_KIND_PAYLOAD2
    = (own head, shared certificate span), so the 9 GB of signed bytes of the round are 10,000 heads + 7 certificate blobs.
    Returns the tuples + the map from every ROUND_CHANGE message to the tuples its validity depends on."""
    height = 1_000_000
    q = 2 * n // 3 + 1
    priv = co.privkeys(3, n)
    addr = co.addresses(priv, N_THREADS)
    rng = np.random.default_rng(4)
    raw = rng.integers(0, 256, 1024, dtype=np.uint8).tobytes()
    ph = proposal_hash(raw, 0)
    view0, view1 = ip.View(height, 0), ip.View(height, 1)
    # round-0 PREPREPARE of validator 0 and PREPAREs of validators 1..n-1
    pp = ip.IbftMessage(view0, bytes(addr[0]), b"", ip.PREPREPARE, ip.PrePrepareMessage(ip.Proposal(raw, 0), ph, None))
    pp_payload = pp.payload_no_sig()
    pp.signature = bytes(co.sign_derived_batch(priv[:1], np.frombuffer(co.keccak256(pp_payload), np.uint8).reshape(1, 32), 1)[0])
    prep_payloads = [ip.IbftMessage(view0, bytes(addr[i]), b"", ip.PREPARE, ip.PrepareMessage(ph)).payload_no_sig() for i in range(1, n)]
    prep_dig = np.stack([np.frombuffer(co.keccak256(p), np.uint8) for p in prep_payloads])
    prep_sigs = co.sign_derived_batch(priv[1:], prep_dig, N_THREADS)
    prepares = [ip.IbftMessage(view0, bytes(addr[i]), bytes(prep_sigs[i - 1]), ip.PREPARE, ip.PrepareMessage(ph)) for i in range(1, n)]
    span = q - 1
    step = max(1, (n - 1 - span) // 2)
    windows = [(w * step, w * step + span) for w in range(3)]          # windows of PREPARE indices (0-based into `prepares`)
    pcs = [ip.PreparedCertificate(pp, prepares[a:b]) for a, b in windows]
    # corrupted variants: one nested PREPARE signature flipped, per certificate
    bad_pos = [windows[w][0] + (7 + 11 * w) % span for w in range(3)]   # index into `prepares`
    bad_sigs = []
    bad_pcs = []
    for w in range(3):
        sig = bytearray(prepares[bad_pos[w]].signature)
        sig[40] ^= 1
        bad_sigs.append(bytes(sig))
        m = prepares[bad_pos[w]]
        bad_msg = ip.IbftMessage(m.view, m.from_, bytes(sig), m.type, m.payload)
        a, b = windows[w]
        msgs = list(prepares[a:b])
        msgs[bad_pos[w] - a] = bad_msg
        bad_pcs.append(ip.PreparedCertificate(pp, msgs))
    short_pc = ip.PreparedCertificate(pp, prepares[: span // 2])     # below quorum
    variants = pcs + bad_pcs + [short_pc]                              # certificate id 0..6
    blobs = [ip.encode_pc(pc) for pc in variants]
    bad_set = set(int(x) for x in rng.choice(np.arange(12, n), size=n_bad, replace=False)) if n_bad else set()
    cert_of = np.zeros(n, np.int32)
    for i in range(n):
        cert_of[i] = i % 3
        if i in bad_set:
            cert_of[i] = 3 + (i % 3)
    cert_of[11] = 6
    cert_of[5] = -1                                                    # no certificate
    # arena: certificate blobs first, then the heads (PayloadNoSig minus the certificate bytes) of every ROUND_CHANGE
    arena = bytearray()
    blob_off = []
    for b in blobs:
        blob_off.append(len(arena))
        arena.extend(b)
    blob_digest_state = {}
    head_off, head_len = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    heads = []
    for i in range(n):
        c = int(cert_of[i])
        if c >= 0:
            head = rc_head(view1, bytes(addr[i]), ip.Proposal(raw, 0), len(blobs[c]))
        else:
            head = ip.IbftMessage(view1, bytes(addr[i]), b"", ip.ROUND_CHANGE, ip.RoundChangeMessage(None, None)).payload_no_sig()
        head_off[i], head_len[i] = len(arena), len(head)
        arena.extend(head)
        heads.append(head)
    # sender signatures of the ROUND_CHANGE messages: digest over head || certificate (9 GB of sponge input at n = 10,000)
    cat = bytearray()
    c_offs, c_lens = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
    rc_digest = np.zeros((n, 32), np.uint8)
    CH = 256                                           # messages hashed per oracle call (bounds the scratch to ~230 MB)
    for lo in range(0, n, CH):
        cat = bytearray()
        offs_l, lens_l = [], []
        for i in range(lo, min(n, lo + CH)):
            c = int(cert_of[i])
            offs_l.append(len(cat))
            cat.extend(heads[i])
            if c >= 0:
                cat.extend(blobs[c])
            lens_l.append(len(cat) - offs_l[-1])
        rc_digest[lo:lo + len(offs_l)] = co.keccak256_batch(bytes(cat), offs_l, lens_l, N_THREADS)
    rc_sigs = co.sign_derived_batch(priv, rc_digest, N_THREADS)
    # ---- the unique tuples
    # [0, n)                 ROUND_CHANGE sender signatures (KIND_PAYLOAD2, or KIND_PAYLOAD for the one without a certificate)
    # n                      the round-0 PREPREPARE
    # [n+1, n+1+(n-1))       the round-0 PREPAREs of validators 1..n-1 (those no window covers are simply never referenced)
    # then                   the three corrupted PREPARE variants
    n_items = n + 1 + (n - 1) + 3
    items = _fill_items(n_items)
    _put_sigs(items[:n], rc_sigs, addr)
    for i in range(n):
        c = int(cert_of[i])
        items["payload_off"][i], items["payload_len"][i] = head_off[i], head_len[i]
        if c >= 0:
            items["kind"][i] = KIND_PAYLOAD2
            items["digest"][i][:8] = np.frombuffer(int(blob_off[c]).to_bytes(8, "little"), np.uint8)
            items["digest"][i][8:12] = np.frombuffer(len(blobs[c]).to_bytes(4, "little"), np.uint8)
        else:
            items["kind"][i] = KIND_PAYLOAD
    # nested messages: plain KIND_PAYLOAD tuples, payload bytes appended to the arena
    def add_payload(idx, payload, sig, frm):
        items["payload_off"][idx], items["payload_len"][idx] = len(arena), len(payload)
        arena.extend(payload)
        items["kind"][idx] = KIND_PAYLOAD
        _put_sigs(items[idx:idx + 1], np.frombuffer(sig, np.uint8).reshape(1, 65), np.frombuffer(frm, np.uint8).reshape(1, 20))
    add_payload(n, pp_payload, pp.signature, bytes(addr[0]))
    for j in range(n - 1):
        add_payload(n + 1 + j, prep_payloads[j], bytes(prep_sigs[j]), bytes(addr[j + 1]))
    for w in range(3):
        add_payload(n + 1 + (n - 1) + w, prep_payloads[bad_pos[w]], bad_sigs[w], bytes(addr[bad_pos[w] + 1]))
    items["group"] = 0
    # per certificate variant: the tuple indices its validity depends on (PREPREPARE + its PREPAREs)
    def cert_tuples(c):
        if c < 3:
            a, b = windows[c]
            return [n] + [n + 1 + j for j in range(a, b)]
        if c < 6:
            a, b = windows[c - 3]
            return [n] + [n + 1 + j if j != bad_pos[c - 3] else n + 1 + (n - 1) + (c - 3) for j in range(a, b)]
        return [n] + [n + 1 + j for j in range(0, span // 2)]
    return dict(items=items, arena=bytes(arena), addrs=addr, n=n, quorum=q, height=height, cert_of=cert_of,
                cert_tuples=[cert_tuples(c) for c in range(7)], cert_sizes=[1 + len(v.prepare_messages) for v in variants],
                raw_proposal=raw, proposal_hash=ph, blobs=blobs, heads=heads, rc_sigs=rc_sigs, variants=variants)


def rc_head(view, frm: bytes, proposal, pc_len: int) -> bytes:
    """PayloadNoSig of a ROUND_CHANGE message MINUS the bytes of its prepared certificate: the certificate is the last field
    (latestPreparedCertificate = 2) of the last field (roundChangeData = 8) of the message, so the signed bytes are exactly
    head || encode_pc(certificate)  (messages/proto/messages.proto:24-44, :75-83; helper.go:13-27)."""
    body_prefix = ip._f_msg(1, ip.encode_proposal(proposal)) + ip._varint((2 << 3) | 2) + ip._varint(pc_len)
    return (ip._f_msg(1, ip.encode_view(view)) + ip._f_bytes(2, frm) + ip._f_varint(4, ip.ROUND_CHANGE) + ip._varint((8 << 3) | 2)
            + ip._varint(len(body_prefix) + pc_len) + body_prefix)


def config4_expected(w, bitmap):
    """validPC-level restatement over the tuple verdicts (core/ibft.go:1162-1231 for the parts a signature decides; the
    structural rules are fixed by construction: same height/round/hash, unique senders, proposer = validator 0):
    ROUND_CHANGE i is valid  <=>  its sender signature is valid AND (no certificate, or: the certificate has quorum size
    (1 + PREPAREs >= Q) and every nested signature is valid).  Returns (valid[n], has_quorum)."""
    bits = np.unpackbits(np.asarray(bitmap, dtype=np.uint32).view(np.uint8), bitorder="little")
    n = w["n"]
    cert_ok = [w["cert_sizes"][c] >= w["quorum"] and all(bits[t] for t in w["cert_tuples"][c]) for c in range(7)]
    valid = np.array([bool(bits[i]) and (w["cert_of"][i] < 0 or cert_ok[int(w["cert_of"][i])]) for i in range(n)])
    return valid, int(valid.sum()) >= w["quorum"]


# ---- loader used by the tests: regenerate (or reuse a local cache whose fingerprint matches the committed pin)
def _sha(a) -> bytes:
    import hashlib
    return hashlib.sha256(a if isinstance(a, (bytes, bytearray)) else np.ascontiguousarray(a).tobytes()).digest()


_HERE = os.path.dirname(os.path.abspath(__file__))
_FULL = {}


def load_full(name: str):
    """name: "config5" or "config4_n10k".  Returns (workload dict, pin npz).  The workload is regenerated with the oracle's bulk
    generators, or read from tests/golden/_cache/ (git-ignored, rebuilt whenever its fingerprint differs from the pin); its
    SHA-256 fingerprints are asserted against the committed pin either way."""
    import pickle
    if name in _FULL:
        return _FULL[name]
    pin = np.load(os.path.join(_HERE, "golden", name + "_pin.npz"))
    cache = os.path.join(_HERE, "golden", "_cache", name + ".pkl")
    w = None
    if os.path.exists(cache):
        try:
            with open(cache, "rb") as f:
                w = pickle.load(f)
            if _sha(w["items"]) != bytes(pin["sha_items"]) or _sha(w["arena"]) != bytes(pin["sha_arena"]):
                w = None
        except Exception:
            w = None
    if w is None:
        full = config5_full() if name == "config5" else config4_n10k()
        keep = ("items", "arena", "tables", "powers", "heights", "n_groups", "group_table", "n_messages", "addrs", "n", "quorum",
                "height", "cert_of", "cert_tuples", "cert_sizes")
        w = {k: v for k, v in full.items() if k in keep}
        try:
            os.makedirs(os.path.dirname(cache), exist_ok=True)
            with open(cache, "wb") as f:
                pickle.dump(w, f, protocol=4)
        except OSError:
            pass
    assert _sha(w["items"]) == bytes(pin["sha_items"]), name + ": regenerated tuples differ from the committed fingerprint"
    assert _sha(w["arena"]) == bytes(pin["sha_arena"]), name + ": regenerated payload arena differs from the committed fingerprint"
    _FULL[name] = (w, pin)
    return _FULL[name]
