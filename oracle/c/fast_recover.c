/* fast_recover.c -- a TUNED CPU arm for bench.py's cpu_baseline / --impl reference: the same recover + compare + membership
 * verdict as oracle_verify_batch (ibft_oracle.c), computed the way a CPU library would (VERDICT r1 "What's weak" 9 / "Next" 7:
 * "a GLV+wNAF path in the C oracle"):
 *     GLV split of both scalars (4 half-scalars of <= 128 bits), width-5 wNAF over a per-signature affine table of odd
 *     multiples of R (one shared inversion), width-11 wNAF over precomputed affine tables of G and lambda*G, ONE interleaved
 *     ladder of <= 129 doublings (2M+5S) with mixed additions (7M+4S), binary (not Fermat) inversions, dedicated squaring.
 * The plain port (ibft_oracle.c: fixed 4-bit windows over 256 doublings, Fermat inversions, no endomorphism) stays what it is --
 * the independent checker of the CUDA path; THIS file is test / bench infrastructure only (oracle/__init__.py), cross-checked
 * bit for bit against the plain port in tests/test_oracle_crypto.py on the third-party vectors, the fixtures and random /
 * adversarial signatures.  Conventions identical to the port: x = r only, v in {0,1}, 1 <= r,s < n, high-s accepted,
 * address = Keccak-256(X||Y)[12:]  (reference call sites: core/backend.go:41-55).
 * The GLV constants are the published ones of the curve (lambda, beta, the lattice basis and its 2^384-scaled reciprocals, as in
 * Gallant-Lambert-Vanstone 2001 / libsecp256k1's scalar_split_lambda); checked numerically in the test file. */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ibft_verify.h"

void oracle_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);
int oracle_item_digest(const ibft_sig_item* it, const uint8_t* arena, size_t arena_len, uint8_t z[32]);

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef struct { u64 l[4]; } w256; /* little-endian limbs */

static const w256 FP = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
static const w256 FN = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
#define FP_C 0x1000003D1ULL
/* 2^256 - n (129 bits) */
static const u64 FN_C[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 1ULL};

/* ---------------------------------------------------------------- 256-bit helpers */
static inline int w_is_zero(const w256* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int w_cmp(const w256* a, const w256* b) {
  for (int i = 3; i >= 0; i--)
    if (a->l[i] != b->l[i]) return a->l[i] < b->l[i] ? -1 : 1;
  return 0;
}
static inline u64 w_add(w256* r, const w256* a, const w256* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; r->l[i] = (u64)c; c >>= 64; }
  return (u64)c;
}
static inline u64 w_sub(w256* r, const w256* a, const w256* b) {
  u64 br = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - br;
    r->l[i] = (u64)d;
    br = (u64)(d >> 64) & 1;
  }
  return br;
}
static inline void w_shr1(w256* a, u64 top) {
  a->l[0] = (a->l[0] >> 1) | (a->l[1] << 63);
  a->l[1] = (a->l[1] >> 1) | (a->l[2] << 63);
  a->l[2] = (a->l[2] >> 1) | (a->l[3] << 63);
  a->l[3] = (a->l[3] >> 1) | (top << 63);
}
static w256 w_from_be(const uint8_t b[32]) {
  w256 r;
  for (int i = 0; i < 4; i++) {
    u64 v = 0;
    for (int j = 0; j < 8; j++) v = (v << 8) | b[8 * (3 - i) + j];
    r.l[i] = v;
  }
  return r;
}
static void w_to_be(const w256* a, uint8_t b[32]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) b[8 * (3 - i) + j] = (uint8_t)(a->l[i] >> (8 * (7 - j)));
}

/* inverse of a (!= 0) modulo the odd prime m: binary extended Euclid, invariants x1*a = u, x2*a = v (mod m) */
static w256 inv_mod(const w256* a, const w256* m) {
  w256 u = *a, v = *m, x1 = {{1, 0, 0, 0}}, x2 = {{0, 0, 0, 0}};
  const w256 one = {{1, 0, 0, 0}};
  while (w_cmp(&u, &one) != 0 && w_cmp(&v, &one) != 0) {
    while (!(u.l[0] & 1)) {
      w_shr1(&u, 0);
      if (x1.l[0] & 1) { u64 c = w_add(&x1, &x1, m); w_shr1(&x1, c); } else w_shr1(&x1, 0);
    }
    while (!(v.l[0] & 1)) {
      w_shr1(&v, 0);
      if (x2.l[0] & 1) { u64 c = w_add(&x2, &x2, m); w_shr1(&x2, c); } else w_shr1(&x2, 0);
    }
    if (w_cmp(&u, &v) >= 0) {
      w_sub(&u, &u, &v);
      if (w_sub(&x1, &x1, &x2)) w_add(&x1, &x1, m);
    } else {
      w_sub(&v, &v, &u);
      if (w_sub(&x2, &x2, &x1)) w_add(&x2, &x2, m);
    }
  }
  return w_cmp(&u, &one) == 0 ? x1 : x2;
}

/* ---------------------------------------------------------------- field F_p.  Values are kept LAZILY reduced: any representative in
 * [0, 2^256) (so either v or v + p for the 2^32+977 smallest residues); additions and subtractions fold the carry / borrow with
 * 2^256 = C (mod p) and never compare with p.  fe_norm makes a value canonical; comparisons go through fe_is_zero / fe_norm. */
static const w256 FP_CW = {{FP_C, 0, 0, 0}};
static inline void fe_reduce(w256* r, const u64 t[8]) {
  /* t = lo + hi * 2^256 = lo + hi * C (mod p) */
  u64 s[5];
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)t[4 + i] * FP_C + t[i]; s[i] = (u64)c; c >>= 64; }
  s[4] = (u64)c; /* < 2^34 */
  c = (u128)s[4] * FP_C + s[0];
  r->l[0] = (u64)c; c >>= 64;
  c += s[1]; r->l[1] = (u64)c; c >>= 64;
  c += s[2]; r->l[2] = (u64)c; c >>= 64;
  c += s[3]; r->l[3] = (u64)c; c >>= 64;
  if (c) { /* one more wrap: add C (cannot carry out again) */
    u128 d = (u128)r->l[0] + FP_C;
    r->l[0] = (u64)d; d >>= 64;
    for (int i = 1; i < 4 && d; i++) { d += r->l[i]; r->l[i] = (u64)d; d >>= 64; }
  }
}
static inline void fe_norm(w256* r) {
  if (w_cmp(r, &FP) >= 0) w_sub(r, r, &FP);
}
static inline int fe_is_zero(const w256* a) { return w_is_zero(a) || w_cmp(a, &FP) == 0; }
static inline void fe_mul(w256* r, const w256* a, const w256* b) {
  u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->l[i] * b->l[j] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
    t[i + 4] = (u64)c;
  }
  fe_reduce(r, t);
}
static inline void fe_sqr(w256* r, const w256* a) {
  /* off-diagonal products once, doubled, plus the squares */
  u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 3; i++) {
    u128 c = 0;
    for (int j = i + 1; j < 4; j++) { c += (u128)a->l[i] * a->l[j] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
    t[i + 4] = (u64)c;
  }
  u64 top = 0;
  for (int i = 0; i < 8; i++) { u64 nt = t[i] >> 63; t[i] = (t[i] << 1) | top; top = nt; }
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    u128 sq = (u128)a->l[i] * a->l[i];
    c += (u128)t[2 * i] + (u64)sq; t[2 * i] = (u64)c; c >>= 64;
    c += (u128)t[2 * i + 1] + (u64)(sq >> 64); t[2 * i + 1] = (u64)c; c >>= 64;
  }
  fe_reduce(r, t);
}
static inline void fe_add(w256* r, const w256* a, const w256* b) {
  u64 c = w_add(r, a, b);
  while (c) c = w_add(r, r, &FP_CW); /* 2^256 = C; a second wrap needs both inputs within C of 2^256 */
}
static inline void fe_sub(w256* r, const w256* a, const w256* b) {
  u64 br = w_sub(r, a, b);
  while (br) br = w_sub(r, r, &FP_CW); /* a - b + 2^256 = a - b + C */
}
static inline void fe_neg(w256* r, const w256* a) {
  const w256 zero = {{0, 0, 0, 0}};
  fe_sub(r, &zero, a);
}
static inline void fe_dbl(w256* r, const w256* a) { fe_add(r, a, a); }
static void fe_sqrn(w256* r, const w256* a, int n) {
  *r = *a;
  for (int i = 0; i < n; i++) fe_sqr(r, r);
}
/* a^((p+1)/4); (p+1)/4 = [223 ones][0][22 ones][0000][11][00] -- the published addition chain for this prime */
static int fe_sqrt(w256* r, const w256* a) {
  w256 x2, x3, x6, x9, x11, x22, x44, x88, x176, x220, x223, t;
  fe_sqr(&t, a); fe_mul(&x2, &t, a);
  fe_sqr(&t, &x2); fe_mul(&x3, &t, a);
  fe_sqrn(&t, &x3, 3); fe_mul(&x6, &t, &x3);
  fe_sqrn(&t, &x6, 3); fe_mul(&x9, &t, &x3);
  fe_sqrn(&t, &x9, 2); fe_mul(&x11, &t, &x2);
  fe_sqrn(&t, &x11, 11); fe_mul(&x22, &t, &x11);
  fe_sqrn(&t, &x22, 22); fe_mul(&x44, &t, &x22);
  fe_sqrn(&t, &x44, 44); fe_mul(&x88, &t, &x44);
  fe_sqrn(&t, &x88, 88); fe_mul(&x176, &t, &x88);
  fe_sqrn(&t, &x176, 44); fe_mul(&x220, &t, &x44);
  fe_sqrn(&t, &x220, 3); fe_mul(&x223, &t, &x3);
  fe_sqrn(&t, &x223, 23); fe_mul(&t, &t, &x22);
  fe_sqrn(&t, &t, 6); fe_mul(&t, &t, &x2);
  fe_sqrn(&t, &t, 2);
  w256 chk, an = *a;
  fe_sqr(&chk, &t);
  fe_norm(&chk);
  fe_norm(&an);
  fe_norm(&t);
  *r = t;
  return w_cmp(&chk, &an) == 0;
}

/* ---------------------------------------------------------------- scalars mod n */
static void sc_reduce512(w256* r, const u64 t[8]) {
  /* fold the high half three times with 2^256 = C_n (mod n), C_n < 2^129 */
  u64 a[8];
  memcpy(a, t, sizeof a);
  for (int round = 0; round < 3; round++) {
    u64 hi[4] = {a[4], a[5], a[6], a[7]};
    if ((hi[0] | hi[1] | hi[2] | hi[3]) == 0) break;
    u64 prod[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 3; j++) { c += (u128)hi[i] * FN_C[j] + prod[i + j]; prod[i + j] = (u64)c; c >>= 64; }
      prod[i + 3] = (u64)c;
    }
    u128 c = 0;
    for (int i = 0; i < 8; i++) { c += (u128)(i < 4 ? a[i] : 0) + prod[i]; a[i] = (u64)c; c >>= 64; }
  }
  w256 v = {{a[0], a[1], a[2], a[3]}};
  /* a[4..7] are zero now (third fold: hi < 2^3) -- up to two subtractions of n remain */
  while (a[4] || w_cmp(&v, &FN) >= 0) {
    u64 br = w_sub(&v, &v, &FN);
    a[4] -= br;
  }
  *r = v;
}
static void sc_mul(w256* r, const w256* a, const w256* b) {
  u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->l[i] * b->l[j] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
    t[i + 4] = (u64)c;
  }
  sc_reduce512(r, t);
}
static void sc_add(w256* r, const w256* a, const w256* b) {
  u64 c = w_add(r, a, b);
  if (c || w_cmp(r, &FN) >= 0) w_sub(r, r, &FN);
}
static void sc_neg(w256* r, const w256* a) {
  if (w_is_zero(a)) { *r = *a; return; }
  w_sub(r, &FN, a);
}

/* k = k1 + k2*lambda (mod n) with |k1|, |k2| <= 2^128: magnitudes + signs */
typedef struct { w256 mag; int neg; } half_t;
static const w256 GLV_G1 = {{0xE893209A45DBB031ULL, 0x3DAA8A1471E8CA7FULL, 0xE86C90E49284EB15ULL, 0x3086D221A7D46BCDULL}};
static const w256 GLV_G2 = {{0x1571B4AE8AC47F71ULL, 0x221208AC9DF506C6ULL, 0x6F547FA90ABFE4C4ULL, 0xE4437ED6010E8828ULL}};
static const w256 GLV_MB1 = {{0x6F547FA90ABFE4C3ULL, 0xE4437ED6010E8828ULL, 0, 0}};
static const w256 GLV_MB2 = {{0xD765CDA83DB1562CULL, 0x8A280AC50774346DULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
static const w256 GLV_LAMBDA = {{0xDF02967C1B23BD72ULL, 0x122E22EA20816678ULL, 0xA5261C028812645AULL, 0x5363AD4CC05C30E0ULL}};
static const w256 FE_BETA = {{0xC1396C28719501EEULL, 0x9CF0497512F58995ULL, 0x6E64479EAC3434E9ULL, 0x7AE96A2B657C0710ULL}};
static const w256 FN_HALF = {{0xDFE92F46681B20A0ULL, 0x5D576E7357A4501DULL, 0xFFFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL}}; /* (n-1)/2 */

static void mul_shift384(w256* r, const w256* a, const w256* b) { /* round(a*b / 2^384) */
  u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->l[i] * b->l[j] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
    t[i + 4] = (u64)c;
  }
  u128 c = (u128)t[6] + (t[5] >> 63);
  r->l[0] = (u64)c; c >>= 64;
  c += t[7]; r->l[1] = (u64)c;
  r->l[2] = (u64)(c >> 64);
  r->l[3] = 0;
}
static void glv_split(const w256* k, half_t* h1, half_t* h2) {
  w256 c1, c2, r1, r2, t;
  mul_shift384(&c1, k, &GLV_G1);
  mul_shift384(&c2, k, &GLV_G2);
  sc_mul(&c1, &c1, &GLV_MB1);
  sc_mul(&c2, &c2, &GLV_MB2);
  sc_add(&r2, &c1, &c2);
  sc_mul(&t, &r2, &GLV_LAMBDA);
  sc_neg(&t, &t);
  sc_add(&r1, k, &t);
  h1->neg = w_cmp(&r1, &FN_HALF) > 0;
  if (h1->neg) sc_neg(&h1->mag, &r1); else h1->mag = r1;
  h2->neg = w_cmp(&r2, &FN_HALF) > 0;
  if (h2->neg) sc_neg(&h2->mag, &r2); else h2->mag = r2;
}

/* width-w NAF of a magnitude < 2^129: digits odd in (-2^(w-1), 2^(w-1)), at most 130 of them; returns the length */
static int wnaf(int8_t* out /* [132] */, int16_t* out16, const w256* mag, int w) {
  u64 k[3] = {mag->l[0], mag->l[1], mag->l[2]};
  int len = 0, pos = 0;
  while (k[0] | k[1] | k[2]) {
    int d = 0;
    if (k[0] & 1) {
      d = (int)(k[0] & ((1u << w) - 1));
      if (d >= (1 << (w - 1))) d -= (1 << w);
      /* k -= d */
      if (d >= 0) {
        u64 b = (u64)d, br = k[0] < b;
        k[0] -= b;
        if (br) { br = k[1] == 0; k[1]--; if (br) k[2]--; }
      } else {
        u64 a = (u64)(-d), c = 0;
        k[0] += a; c = k[0] < a;
        if (c) { k[1]++; if (k[1] == 0) k[2]++; }
      }
    }
    if (out) out[pos] = (int8_t)d; else out16[pos] = (int16_t)d;
    if (d) len = pos + 1;
    pos++;
    k[0] = (k[0] >> 1) | (k[1] << 63);
    k[1] = (k[1] >> 1) | (k[2] << 63);
    k[2] >>= 1;
  }
  for (int i = pos; i < 132; i++) { if (out) out[i] = 0; else out16[i] = 0; }
  return len;
}

/* ---------------------------------------------------------------- points */
typedef struct { w256 x, y; } aff_t;
typedef struct { w256 x, y, z; int inf; } jac_t;

static void jac_double(jac_t* r, const jac_t* p) { /* dbl-2009-l, a = 0: 2M + 5S */
  if (p->inf || fe_is_zero(&p->y)) { r->inf = 1; return; }
  w256 A, B, C, D, E, F, t, z3;
  fe_sqr(&A, &p->x);
  fe_sqr(&B, &p->y);
  fe_sqr(&C, &B);
  fe_add(&t, &p->x, &B); fe_sqr(&t, &t); fe_sub(&t, &t, &A); fe_sub(&t, &t, &C); fe_dbl(&D, &t);
  fe_dbl(&E, &A); fe_add(&E, &E, &A);
  fe_sqr(&F, &E);
  fe_mul(&z3, &p->y, &p->z); fe_dbl(&z3, &z3);
  fe_dbl(&t, &D); fe_sub(&r->x, &F, &t);
  fe_sub(&t, &D, &r->x); fe_mul(&t, &E, &t);
  fe_dbl(&C, &C); fe_dbl(&C, &C); fe_dbl(&C, &C);
  fe_sub(&r->y, &t, &C);
  r->z = z3;
  r->inf = 0;
}
static void jac_add_aff(jac_t* r, const jac_t* p, const w256* qx, const w256* qy) { /* mixed addition, exceptional cases exact */
  if (p->inf) { r->x = *qx; r->y = *qy; r->z = (w256){{1, 0, 0, 0}}; r->inf = 0; return; }
  w256 z2, u2, s2, h, rr, h2, h3, v, t;
  fe_sqr(&z2, &p->z);
  fe_mul(&u2, qx, &z2);
  fe_mul(&s2, qy, &z2); fe_mul(&s2, &s2, &p->z);
  fe_sub(&h, &u2, &p->x);
  fe_sub(&rr, &s2, &p->y);
  if (fe_is_zero(&h)) {
    if (fe_is_zero(&rr)) { jac_t d = *p; jac_double(r, &d); return; }
    r->inf = 1;
    return;
  }
  fe_sqr(&h2, &h);
  fe_mul(&h3, &h2, &h);
  fe_mul(&v, &p->x, &h2);
  fe_sqr(&t, &rr); fe_sub(&t, &t, &h3); fe_sub(&t, &t, &v); fe_sub(&t, &t, &v);
  w256 x3 = t, y3;
  fe_sub(&t, &v, &x3); fe_mul(&t, &t, &rr);
  fe_mul(&y3, &p->y, &h3); fe_sub(&y3, &t, &y3);
  fe_mul(&r->z, &p->z, &h);
  r->x = x3; r->y = y3; r->inf = 0;
}
static int jac_to_aff(aff_t* r, const jac_t* p) {
  if (p->inf || fe_is_zero(&p->z)) return 0;
  w256 zn = p->z;
  fe_norm(&zn);
  w256 zi = inv_mod(&zn, &FP), zi2, zi3;
  fe_sqr(&zi2, &zi);
  fe_mul(&zi3, &zi2, &zi);
  fe_mul(&r->x, &p->x, &zi2);
  fe_mul(&r->y, &p->y, &zi3);
  fe_norm(&r->x);
  fe_norm(&r->y);
  return 1;
}

/* odd multiples 1, 3, ..., 2*cnt-1 of P as AFFINE points.  The chain P, P+2P, ... runs on the curve ISOMORPHIC to this one by
 * (x, y) -> (x Z^2, y Z^3) with Z the Jacobian z of 2P, where 2P is affine -- so every step is a mixed addition and no inversion
 * is needed for 2P; a point (x', y', z') of that curve is (x', y', z' Z) here.  One shared inversion (Montgomery's trick) makes the
 * table affine.  J, acc: caller's scratch of cnt elements. */
static void odd_multiples(aff_t* out, int cnt, const aff_t* P, jac_t* J, w256* acc) {
  jac_t p1 = {P->x, P->y, {{1, 0, 0, 0}}, 0}, d2;
  jac_double(&d2, &p1);
  w256 z2, z3;
  fe_sqr(&z2, &d2.z);
  fe_mul(&z3, &z2, &d2.z);
  fe_mul(&J[0].x, &P->x, &z2);
  fe_mul(&J[0].y, &P->y, &z3);
  J[0].z = (w256){{1, 0, 0, 0}};
  J[0].inf = 0;
  for (int i = 1; i < cnt; i++) jac_add_aff(&J[i], &J[i - 1], &d2.x, &d2.y);
  for (int i = 0; i < cnt; i++) fe_mul(&J[i].z, &J[i].z, &d2.z); /* back to this curve */
  acc[0] = J[0].z;
  for (int i = 1; i < cnt; i++) fe_mul(&acc[i], &acc[i - 1], &J[i].z);
  w256 last = acc[cnt - 1];
  fe_norm(&last);
  w256 inv = inv_mod(&last, &FP);
  for (int i = cnt - 1; i >= 0; i--) {
    w256 zi, zi2, zi3;
    if (i) { fe_mul(&zi, &inv, &acc[i - 1]); fe_mul(&inv, &inv, &J[i].z); } else zi = inv;
    fe_sqr(&zi2, &zi);
    fe_mul(&zi3, &zi2, &zi);
    fe_mul(&out[i].x, &J[i].x, &zi2);
    fe_mul(&out[i].y, &J[i].y, &zi3);
  }
}

#define WG 11            /* generator window */
#define G_ENTRIES (1 << (WG - 2))
#define WR 5             /* per-signature window */
#define R_ENTRIES (1 << (WR - 2))
static aff_t G_TAB[G_ENTRIES], GL_TAB[G_ENTRIES]; /* odd multiples of G and of lambda*G */
static pthread_once_t g_fast_once = PTHREAD_ONCE_INIT;
static void fast_init(void) {
  const aff_t G = {{{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}},
                   {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}}};
  static jac_t J[G_ENTRIES];
  static w256 acc[G_ENTRIES];
  odd_multiples(G_TAB, G_ENTRIES, &G, J, acc);
  for (int i = 0; i < G_ENTRIES; i++) { fe_mul(&GL_TAB[i].x, &G_TAB[i].x, &FE_BETA); GL_TAB[i].y = G_TAB[i].y; }
}

/* u1*G + u2*R */
static void ecmult_fast(jac_t* out, const w256* u1, const w256* u2, const aff_t* R) {
  half_t g1, g2, r1, r2;
  glv_split(u1, &g1, &g2);
  glv_split(u2, &r1, &r2);
  int16_t ng1[132], ng2[132];
  int8_t nr1[132], nr2[132];
  int len = wnaf(NULL, ng1, &g1.mag, WG), l;
  if ((l = wnaf(NULL, ng2, &g2.mag, WG)) > len) len = l;
  if ((l = wnaf(nr1, NULL, &r1.mag, WR)) > len) len = l;
  if ((l = wnaf(nr2, NULL, &r2.mag, WR)) > len) len = l;
  aff_t RT[R_ENTRIES], RLT[R_ENTRIES];
  jac_t J[R_ENTRIES];
  w256 accz[R_ENTRIES];
  odd_multiples(RT, R_ENTRIES, R, J, accz);
  for (int i = 0; i < R_ENTRIES; i++) { fe_mul(&RLT[i].x, &RT[i].x, &FE_BETA); RLT[i].y = RT[i].y; }
  jac_t acc;
  acc.inf = 1;
  for (int i = len - 1; i >= 0; i--) {
    if (!acc.inf) { jac_t t = acc; jac_double(&acc, &t); }
    const int dd[4] = {nr1[i], nr2[i], ng1[i], ng2[i]};
    const int neg[4] = {r1.neg, r2.neg, g1.neg, g2.neg};
    const aff_t* tab[4] = {RT, RLT, G_TAB, GL_TAB};
    for (int s = 0; s < 4; s++) {
      int d = dd[s];
      if (!d) continue;
      int ng = (d < 0) != (neg[s] != 0);
      const aff_t* e = &tab[s][((d < 0 ? -d : d) - 1) >> 1];
      w256 y = e->y;
      if (ng) fe_neg(&y, &y);
      jac_t t = acc;
      jac_add_aff(&acc, &t, &e->x, &y);
    }
  }
  *out = acc;
}

/* 1 = an address was recovered into addr */
int fast_ecrecover_address(const uint8_t digest[32], const uint8_t r_be[32], const uint8_t s_be[32], uint8_t v, uint8_t addr[20]) {
  pthread_once(&g_fast_once, fast_init);
  if (v > 1) return 0;
  w256 r = w_from_be(r_be), s = w_from_be(s_be), z = w_from_be(digest);
  if (w_is_zero(&r) || w_is_zero(&s) || w_cmp(&r, &FN) >= 0 || w_cmp(&s, &FN) >= 0) return 0;
  if (w_cmp(&z, &FN) >= 0) w_sub(&z, &z, &FN);
  /* R = lift_x(r, v): y^2 = x^3 + 7 (r < n < p: a valid field element) */
  aff_t R;
  R.x = r;
  w256 t, seven = {{7, 0, 0, 0}};
  fe_sqr(&t, &r); fe_mul(&t, &t, &r); fe_add(&t, &t, &seven);
  if (!fe_sqrt(&R.y, &t)) return 0;
  if ((R.y.l[0] & 1) != v) fe_neg(&R.y, &R.y);
  w256 rinv = inv_mod(&r, &FN), u1, u2;
  sc_mul(&u1, &z, &rinv); sc_neg(&u1, &u1);
  sc_mul(&u2, &s, &rinv);
  jac_t Q;
  ecmult_fast(&Q, &u1, &u2, &R);
  aff_t A;
  if (!jac_to_aff(&A, &Q)) return 0;
  uint8_t xy[64], h[32];
  w_to_be(&A.x, xy);
  w_to_be(&A.y, xy + 32);
  oracle_keccak256(xy, 64, h);
  memcpy(addr, h + 12, 20);
  return 1;
}

static int cmp20f(const void* a, const void* b) { return memcmp(a, b, 20); }
typedef struct {
  const ibft_sig_item* items; uint32_t lo, hi; const uint8_t* arena; size_t arena_len;
  const uint8_t* table; uint32_t table_n; uint8_t* verdict;
} fjob;
static void* fworker(void* a) {
  fjob* j = (fjob*)a;
  for (uint32_t i = j->lo; i < j->hi; i++) {
    const ibft_sig_item* it = &j->items[i];
    uint8_t z[32], addr[20];
    int ok = oracle_item_digest(it, j->arena, j->arena_len, z) && fast_ecrecover_address(z, it->r, it->s, it->v, addr) &&
             memcmp(addr, it->signer, 20) == 0;
    if (ok && j->table) ok = bsearch(addr, j->table, j->table_n, 20, cmp20f) != NULL;
    j->verdict[i] = (uint8_t)ok;
  }
  return NULL;
}
/* one validator table (may be NULL) for all items; bitmap gets bit i%32 of word i/32 */
int fast_verify_batch(const ibft_sig_item* items, uint32_t n, const uint8_t* arena, size_t arena_len, const uint8_t* table,
                      uint32_t table_n, int n_threads, uint32_t* bitmap) {
  pthread_once(&g_fast_once, fast_init);
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  uint8_t* verdict = (uint8_t*)calloc(n ? n : 1, 1);
  uint8_t* sorted = NULL;
  if (!verdict) return -1;
  if (table) {
    sorted = (uint8_t*)malloc((size_t)table_n * 20 + 1);
    if (!sorted) { free(verdict); return -1; }
    memcpy(sorted, table, (size_t)table_n * 20);
    qsort(sorted, table_n, 20, cmp20f);
  }
  pthread_t th[256];
  fjob jobs[256];
  uint32_t per = (n + (uint32_t)n_threads - 1) / (uint32_t)n_threads;
  int started = 0;
  for (int k = 0; k < n_threads; k++) {
    uint32_t lo = (uint32_t)k * per, hi = lo + per > n ? n : lo + per;
    if (lo >= hi) break;
    jobs[k] = (fjob){items, lo, hi, arena, arena_len, sorted, table_n, verdict};
    if (n_threads == 1) fworker(&jobs[k]);
    else pthread_create(&th[k], NULL, fworker, &jobs[k]);
    started++;
  }
  if (n_threads > 1)
    for (int k = 0; k < started; k++) pthread_join(th[k], NULL);
  memset(bitmap, 0, ((size_t)n + 31) / 32 * 4);
  for (uint32_t i = 0; i < n; i++)
    if (verdict[i]) bitmap[i / 32] |= 1u << (i % 32);
  free(verdict);
  free(sorted);
  return 0;
}
