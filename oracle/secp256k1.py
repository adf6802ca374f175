"""secp256k1 ECDSA recover / verify / sign with Python big integers.

Oracle / test infrastructure only (see oracle/__init__.py).  The reference defines only the call
sites (core/backend.go:41-45 IsValidValidator, :53-55 IsValidCommittedSeal; callers
core/ibft.go:735,943,1128,1213,1220); the arithmetic below restates the published algorithms:
SEC 1 v2 §4.1.6 (public-key recovery), §4.1.4 (verification), SEC 2 v2 §2.4.1 (domain
parameters), RFC 6979 §3.2 (deterministic nonces, HMAC-SHA256) for reproducible fixtures.

Conventions ([EXTERNAL] -- fixed once here, mirrored by the C oracle and the CUDA
kernels):
  * signature = 65 bytes R||S||V, R,S big-endian, V in {0,1} is the parity of R.y;
  * valid iff 1 <= r < n, 1 <= s < n, V in {0,1}, x = r is the abscissa of a curve point
    (x = r + n is NOT tried), and the recovered key is not the point at infinity;
  * high-s signatures are ACCEPTED (recover-style APIs do not enforce low-s);
  * address = Keccak-256(X||Y as 64 big-endian bytes)[12:32].
"""
from __future__ import annotations

import hashlib
import hmac

from .keccak import keccak256

P = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
G = (GX, GY)
INF = None  # point at infinity


def on_curve(pt) -> bool:
    if pt is INF:
        return True
    x, y = pt
    return (y * y - x * x * x - 7) % P == 0


def point_add(a, b):
    if a is INF:
        return b
    if b is INF:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def point_neg(a):
    return INF if a is INF else (a[0], (-a[1]) % P)


def point_mul(k: int, pt):
    k %= N
    acc = INF
    while k:
        if k & 1:
            acc = point_add(acc, pt)
        pt = point_add(pt, pt)
        k >>= 1
    return acc


def lift_x(x: int, odd: int):
    """Curve point with abscissa x and y parity `odd`, or None if x^3+7 is a non-residue / x >= p."""
    if x >= P:
        return None
    y2 = (pow(x, 3, P) + 7) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        return None
    if (y & 1) != odd:
        y = P - y
    return (x, y)


def recover_pubkey(z: int, r: int, s: int, v: int):
    """SEC 1 §4.1.6 with j = 0 only: Q = r^-1 (s R - z G).  Returns affine (X, Y) or None."""
    if not (1 <= r < N and 1 <= s < N) or v not in (0, 1):
        return None
    R = lift_x(r, v)
    if R is None:
        return None
    rinv = pow(r, -1, N)
    u1 = (-z * rinv) % N
    u2 = (s * rinv) % N
    Q = point_add(point_mul(u1, G), point_mul(u2, R))
    return Q  # INF (None) when the sum vanishes


def pubkey_to_address(pt) -> bytes:
    x, y = pt
    return keccak256(x.to_bytes(32, "big") + y.to_bytes(32, "big"))[12:]


def privkey_to_pubkey(d: int):
    return point_mul(d, G)


def privkey_to_address(d: int) -> bytes:
    return pubkey_to_address(privkey_to_pubkey(d))


def ecrecover_address(digest: bytes, sig: bytes):
    """digest: 32 bytes; sig: 65 bytes R||S||V.  Returns the 20-byte address or None."""
    if len(digest) != 32 or len(sig) != 65:
        return None
    r = int.from_bytes(sig[0:32], "big")
    s = int.from_bytes(sig[32:64], "big")
    q = recover_pubkey(int.from_bytes(digest, "big"), r, s, sig[64])
    if q is None:
        return None
    return pubkey_to_address(q)


def verify(z: int, r: int, s: int, pub) -> bool:
    """Plain ECDSA verification (SEC 1 §4.1.4) -- used only to cross-check recover."""
    if not (1 <= r < N and 1 <= s < N) or pub is INF:
        return False
    w = pow(s, -1, N)
    pt = point_add(point_mul(z * w % N, G), point_mul(r * w % N, pub))
    return pt is not INF and pt[0] % N == r


def rfc6979_k(d: int, digest: bytes) -> int:
    """RFC 6979 §3.2 with HMAC-SHA256; digest is the 32-byte message hash (already hashed)."""
    x = d.to_bytes(32, "big")
    h1 = (int.from_bytes(digest, "big") % N).to_bytes(32, "big")
    V = b"\x01" * 32
    K = b"\x00" * 32
    K = hmac.new(K, V + b"\x00" + x + h1, hashlib.sha256).digest()
    V = hmac.new(K, V, hashlib.sha256).digest()
    K = hmac.new(K, V + b"\x01" + x + h1, hashlib.sha256).digest()
    V = hmac.new(K, V, hashlib.sha256).digest()
    while True:
        V = hmac.new(K, V, hashlib.sha256).digest()
        k = int.from_bytes(V, "big")
        if 1 <= k < N:
            return k
        K = hmac.new(K, V + b"\x00", hashlib.sha256).digest()
        V = hmac.new(K, V, hashlib.sha256).digest()


def sign(d: int, digest: bytes, low_s: bool = True) -> bytes:
    """Deterministic 65-byte recoverable signature R||S||V over a 32-byte digest."""
    z = int.from_bytes(digest, "big")
    k = rfc6979_k(d, digest)
    while True:
        Rp = point_mul(k, G)
        r = Rp[0] % N
        s = pow(k, -1, N) * (z + r * d) % N
        if r != 0 and s != 0 and Rp[0] < N:
            break
        k = (k + 1) % N or 1  # astronomically unlikely; keeps the function total
    v = Rp[1] & 1
    if low_s and s > N // 2:
        s = N - s
        v ^= 1
    return r.to_bytes(32, "big") + s.to_bytes(32, "big") + bytes([v])
