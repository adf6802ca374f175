"""Keccak-256 (original Keccak padding 0x01, NOT SHA3's 0x06), pure Python.

Oracle / test infrastructure only (see oracle/__init__.py).  Restates the published
Keccak-f[1600] permutation (Keccak reference, rate 1088 bits / capacity 512 for the 256-bit
output).  The reference only names the hash in comments: core/ibft.go:648 ("hash matches
keccak(proposal)") and messages/proto/messages.proto:51,61,67 ("Keccak hash of the proposal").
"""
from __future__ import annotations

RATE = 136
_MASK = (1 << 64) - 1

RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
# rotation offsets r[x][y]
ROT = [
    [0, 36, 3, 41, 18],
    [1, 44, 10, 45, 2],
    [62, 6, 43, 15, 61],
    [28, 55, 25, 21, 56],
    [27, 20, 39, 8, 14],
]


def _rol(v: int, n: int) -> int:
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _MASK if n else v


def keccak_f1600(a: list[int]) -> list[int]:
    """a: 25 lanes, index x + 5*y."""
    for rnd in range(24):
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], ROT[x][y])
        a = [b[i] ^ ((~b[(i % 5 + 1) % 5 + 5 * (i // 5)]) & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] ^= RC[rnd]
    return a


def keccak256(data: bytes) -> bytes:
    data = bytes(data)
    padded = bytearray(data)
    padlen = RATE - (len(data) % RATE)
    padded += b"\x00" * padlen
    padded[len(data)] ^= 0x01
    padded[-1] ^= 0x80
    st = [0] * 25
    for off in range(0, len(padded), RATE):
        blk = padded[off:off + RATE]
        for i in range(RATE // 8):
            st[i] ^= int.from_bytes(blk[8 * i:8 * i + 8], "little")
        st = keccak_f1600(st)
    return b"".join(st[i].to_bytes(8, "little") for i in range(4))
