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
