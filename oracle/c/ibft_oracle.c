/*
 * CPU oracle (plain C) for the go-ibft message-verification hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.  Used by
 * tests/ (as the checker), __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs (as the timed CPU baseline, kind "port").
 *
 * PARITY UNPINNED BY THE REFERENCE: go-ibft ships no cryptography; the functions below restate
 * what the embedder must do behind core.Verifier:
 *   - IsValidValidator      core/backend.go:41-45  (callers core/ibft.go:735,1128,1213,1220)
 *   - IsValidCommittedSeal  core/backend.go:53-55  (caller  core/ibft.go:943)
 *   - IsValidProposalHash   core/backend.go:50-51  (callers core/ibft.go:545,649,781,858,938)
 * using the conventions of SURVEY.md §8(c) [EXTERNAL]: secp256k1 (SEC 2 v2 §2.4.1), public-key
 * recovery SEC 1 v2 §4.1.6 with x = r only, original Keccak-256, 65-byte R||S||V, V in {0,1},
 * address = Keccak-256(X||Y)[12:], high-s accepted.  Pinned by known-answer vectors and by
 * agreement with oracle/secp256k1.py (Python big ints) and the `cryptography` package (OpenSSL);
 * see tests/test_oracle_crypto.py.
 *
 * Deliberately a DIFFERENT algorithm from the CUDA kernels (4x64-bit limbs, Fermat inversion,
 * fixed 4-bit windows, no GLV, no safegcd) so that agreement is evidence, not tautology.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ibft_verify.h"

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------ Keccak-256 */
static const u64 KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

static inline u64 rol64(u64 v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; }

static void keccak_f(u64 a[25]) {
  for (int rnd = 0; rnd < 24; rnd++) {
    u64 c[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) {
      u64 d = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
      for (int y = 0; y < 25; y += 5) a[x + y] ^= d;
    }
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(a[x + 5 * y], KROT[x + 5 * y]);
    for (int y = 0; y < 25; y += 5)
      for (int x = 0; x < 5; x++) a[x + y] = b[x + y] ^ (~b[(x + 1) % 5 + y] & b[(x + 2) % 5 + y]);
    a[0] ^= KRC[rnd];
  }
}

void oracle_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
  u64 st[25];
  uint8_t blk[136];
  memset(st, 0, sizeof st);
  while (len >= 136) {
    for (int i = 0; i < 17; i++) { u64 w; memcpy(&w, data + 8 * i, 8); st[i] ^= w; }
    keccak_f(st);
    data += 136; len -= 136;
  }
  memset(blk, 0, sizeof blk);
  memcpy(blk, data, len);
  blk[len] ^= 0x01;
  blk[135] ^= 0x80;
  for (int i = 0; i < 17; i++) { u64 w; memcpy(&w, blk + 8 * i, 8); st[i] ^= w; }
  keccak_f(st);
  memcpy(out, st, 32);
}

/* ------------------------------------------------------------------ 256-bit helpers (4x64 LE limbs) */
typedef struct { u64 l[4]; } u256;

static const u256 FP = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
static const u256 FN = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
#define FP_C 0x1000003D1ULL /* 2^256 - p */
/* 2^256 - n (129 bits) */
static const u64 FN_C[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 1ULL};

static u256 from_be(const uint8_t b[32]) {
  u256 r;
  for (int i = 0; i < 4; i++) {
    u64 w = 0;
    for (int j = 0; j < 8; j++) w = (w << 8) | b[(3 - i) * 8 + j];
    r.l[i] = w;
  }
  return r;
}
static void to_be(const u256* a, uint8_t b[32]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (uint8_t)(a->l[i] >> (56 - 8 * j));
}
static int cmp256(const u256* a, const u256* b) {
  for (int i = 3; i >= 0; i--) {
    if (a->l[i] < b->l[i]) return -1;
    if (a->l[i] > b->l[i]) return 1;
  }
  return 0;
}
static int is_zero256(const u256* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static u64 add256(u256* r, const u256* a, const u256* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; r->l[i] = (u64)c; c >>= 64; }
  return (u64)c;
}
static u64 sub256(u256* r, const u256* a, const u256* b) {
  u64 brw = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a->l[i] - b->l[i] - brw;
    r->l[i] = (u64)t;
    brw = (u64)(t >> 64) & 1;
  }
  return brw;
}
static void mul512(u64 r[8], const u256* a, const u256* b) {
  memset(r, 0, 8 * sizeof(u64));
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->l[i] * b->l[j] + r[i + j]; r[i + j] = (u64)c; c >>= 64; }
    r[i + 4] = (u64)c;
  }
}

/* ---- field mod p (always fully reduced) */
static u256 fp_add(const u256* a, const u256* b) {
  u256 r, t;
  u64 c = add256(&r, a, b);
  if (c || cmp256(&r, &FP) >= 0) { sub256(&t, &r, &FP); return t; }
  return r;
}
static u256 fp_sub(const u256* a, const u256* b) {
  u256 r, t;
  if (sub256(&r, a, b)) { add256(&t, &r, &FP); return t; }
  return r;
}
static u256 fp_neg(const u256* a) { u256 z = {{0, 0, 0, 0}}; return fp_sub(&z, a); }
static u256 fp_mul(const u256* a, const u256* b) {
  u64 w[8];
  mul512(w, a, b);
  /* fold the high 256 bits: r = lo + hi*C  (fits in 256+34 bits) */
  u64 t[5];
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)w[i + 4] * FP_C + w[i]; t[i] = (u64)c; c >>= 64; }
  t[4] = (u64)c;
  /* fold t[4] */
  u256 r;
  c = (u128)t[4] * FP_C + t[0]; r.l[0] = (u64)c; c >>= 64;
  for (int i = 1; i < 4; i++) { c += t[i]; r.l[i] = (u64)c; c >>= 64; }
  if ((u64)c) { /* wrapped past 2^256 once more: add C (cannot carry again) */
    u128 d = (u128)r.l[0] + FP_C; r.l[0] = (u64)d; d >>= 64;
    for (int i = 1; i < 4; i++) { d += r.l[i]; r.l[i] = (u64)d; d >>= 64; }
  }
  if (cmp256(&r, &FP) >= 0) { u256 s; sub256(&s, &r, &FP); return s; }
  return r;
}
static u256 fp_sqr(const u256* a) { return fp_mul(a, a); }
static u256 fp_sqrn(u256 a, int n) {
  for (int i = 0; i < n; i++) a = fp_sqr(&a);
  return a;
}
/* x^(2^223 - 1) and the two helpers the tails need (standard secp256k1 addition chain: x2,x3,x6,x9,x11,x22,x44,x88,x176,x220,x223) */
static void fp_chain223(const u256* a, u256* x2, u256* x22, u256* x223) {
  u256 t = fp_sqr(a); *x2 = fp_mul(&t, a);
  t = fp_sqr(x2); u256 x3 = fp_mul(&t, a);
  t = fp_sqrn(x3, 3); u256 x6 = fp_mul(&t, &x3);
  t = fp_sqrn(x6, 3); u256 x9 = fp_mul(&t, &x3);
  t = fp_sqrn(x9, 2); u256 x11 = fp_mul(&t, x2);
  t = fp_sqrn(x11, 11); *x22 = fp_mul(&t, &x11);
  t = fp_sqrn(*x22, 22); u256 x44 = fp_mul(&t, x22);
  t = fp_sqrn(x44, 44); u256 x88 = fp_mul(&t, &x44);
  t = fp_sqrn(x88, 88); u256 x176 = fp_mul(&t, &x88);
  t = fp_sqrn(x176, 44); u256 x220 = fp_mul(&t, &x44);
  t = fp_sqrn(x220, 3); *x223 = fp_mul(&t, &x3);
}
static u256 fp_inv(const u256* a) { /* a^(p-2), p-2 = [223 ones][0][22 ones][0000][1][0][11][0][1] */
  u256 x2, x22, x223;
  fp_chain223(a, &x2, &x22, &x223);
  u256 t = fp_sqrn(x223, 23); t = fp_mul(&t, &x22);
  t = fp_sqrn(t, 5); t = fp_mul(&t, a);
  t = fp_sqrn(t, 3); t = fp_mul(&t, &x2);
  t = fp_sqrn(t, 2); t = fp_mul(&t, a);
  return t;
}
static int fp_sqrt(u256* r, const u256* a) { /* a^((p+1)/4), (p+1)/4 = [223 ones][0][22 ones][0000][11][00]; p = 3 mod 4 */
  u256 x2, x22, x223;
  fp_chain223(a, &x2, &x22, &x223);
  u256 t = fp_sqrn(x223, 23); t = fp_mul(&t, &x22);
  t = fp_sqrn(t, 6); t = fp_mul(&t, &x2);
  t = fp_sqrn(t, 2);
  u256 y2 = fp_sqr(&t);
  *r = t;
  return cmp256(&y2, a) == 0;
}

/* ---- scalars mod n */
static void fn_reduce512(u256* r, const u64 w[8]) {
  /* value = lo + hi * (2^256 mod n), iterate until the high part vanishes */
  u64 cur[8];
  memcpy(cur, w, sizeof cur);
  for (int iter = 0; iter < 4; iter++) {
    u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    /* t = hi(4 limbs) * FN_C(3 limbs) */
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 3; j++) { c += (u128)cur[4 + i] * FN_C[j] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
      t[i + 3] += (u64)c; /* cannot overflow: product < 2^(256+129) */
    }
    u128 c = 0;
    for (int i = 0; i < 8; i++) { c += (u128)t[i] + (i < 4 ? cur[i] : 0); cur[i] = (u64)c; c >>= 64; }
  }
  u256 v = {{cur[0], cur[1], cur[2], cur[3]}};
  /* after the folds the high limbs are zero; at most a couple of subtractions remain */
  while (cur[4] | cur[5] | cur[6] | cur[7]) { /* defensive; not expected */
    u64 brw = sub256(&v, &v, &FN);
    cur[4] -= brw;
  }
  while (cmp256(&v, &FN) >= 0) sub256(&v, &v, &FN);
  *r = v;
}
static u256 fn_mul(const u256* a, const u256* b) {
  u64 w[8];
  u256 r;
  mul512(w, a, b);
  fn_reduce512(&r, w);
  return r;
}
static u256 fn_neg(const u256* a) {
  u256 r;
  if (is_zero256(a)) return *a;
  sub256(&r, &FN, a);
  return r;
}
static u256 fn_inv(const u256* a) { /* a^(n-2), fixed 4-bit window */
  u256 e = FN; e.l[0] -= 2;
  u256 tab[16];
  tab[0] = (u256){{1, 0, 0, 0}};
  tab[1] = *a;
  for (int i = 2; i < 16; i++) tab[i] = fn_mul(&tab[i - 1], a);
  u256 r = tab[(e.l[3] >> 60) & 15];
  for (int nib = 62; nib >= 0; nib--) {
    for (int k = 0; k < 4; k++) r = fn_mul(&r, &r);
    unsigned d = (unsigned)(e.l[nib / 16] >> (4 * (nib % 16))) & 15;
    if (d) r = fn_mul(&r, &tab[d]);
  }
  return r;
}

/* ------------------------------------------------------------------ curve (Jacobian, complete via branches) */
typedef struct { u256 x, y, z; int inf; } jac;
typedef struct { u256 x, y; int inf; } aff;

static const aff GEN = {
    {{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}},
    {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}}, 0};

static jac jac_from_aff(const aff* a) {
  jac r; r.x = a->x; r.y = a->y; r.z = (u256){{1, 0, 0, 0}}; r.inf = a->inf;
  return r;
}
static jac jac_double(const jac* p) {
  jac r;
  if (p->inf || is_zero256(&p->y)) { r = *p; r.inf = 1; return r; }
  u256 a = fp_sqr(&p->x), b = fp_sqr(&p->y), c = fp_sqr(&b);
  u256 t = fp_add(&p->x, &b); t = fp_sqr(&t); t = fp_sub(&t, &a); t = fp_sub(&t, &c);
  u256 d = fp_add(&t, &t);
  u256 e = fp_add(&a, &a); e = fp_add(&e, &a);
  u256 f = fp_sqr(&e);
  u256 x3 = fp_sub(&f, &d); x3 = fp_sub(&x3, &d);
  u256 c8 = fp_add(&c, &c); c8 = fp_add(&c8, &c8); c8 = fp_add(&c8, &c8);
  u256 y3 = fp_sub(&d, &x3); y3 = fp_mul(&e, &y3); y3 = fp_sub(&y3, &c8);
  u256 z3 = fp_mul(&p->y, &p->z); z3 = fp_add(&z3, &z3);
  r.x = x3; r.y = y3; r.z = z3; r.inf = 0;
  return r;
}
static jac jac_add(const jac* p, const jac* q) {
  if (p->inf) return *q;
  if (q->inf) return *p;
  u256 z1z1 = fp_sqr(&p->z), z2z2 = fp_sqr(&q->z);
  u256 u1 = fp_mul(&p->x, &z2z2), u2 = fp_mul(&q->x, &z1z1);
  u256 s1 = fp_mul(&p->y, &q->z); s1 = fp_mul(&s1, &z2z2);
  u256 s2 = fp_mul(&q->y, &p->z); s2 = fp_mul(&s2, &z1z1);
  u256 h = fp_sub(&u2, &u1), rr = fp_sub(&s2, &s1);
  if (is_zero256(&h)) {
    if (is_zero256(&rr)) return jac_double(p);
    jac r = *p; r.inf = 1; return r;
  }
  u256 h2 = fp_sqr(&h), h3 = fp_mul(&h, &h2), v = fp_mul(&u1, &h2);
  u256 x3 = fp_sqr(&rr); x3 = fp_sub(&x3, &h3); x3 = fp_sub(&x3, &v); x3 = fp_sub(&x3, &v);
  u256 y3 = fp_sub(&v, &x3); y3 = fp_mul(&rr, &y3);
  u256 t = fp_mul(&s1, &h3); y3 = fp_sub(&y3, &t);
  u256 z3 = fp_mul(&p->z, &q->z); z3 = fp_mul(&z3, &h);
  jac r; r.x = x3; r.y = y3; r.z = z3; r.inf = 0;
  return r;
}
static aff jac_to_aff(const jac* p) {
  aff r;
  if (p->inf) { memset(&r, 0, sizeof r); r.inf = 1; return r; }
  u256 zi = fp_inv(&p->z), zi2 = fp_sqr(&zi), zi3 = fp_mul(&zi2, &zi);
  r.x = fp_mul(&p->x, &zi2); r.y = fp_mul(&p->y, &zi3); r.inf = 0;
  return r;
}

/* 4-bit window tables: tab[i] = (i+1) * P for i in 0..14 */
static void make_table(jac tab[15], const aff* p) {
  tab[0] = jac_from_aff(p);
  for (int i = 1; i < 15; i++) tab[i] = jac_add(&tab[i - 1], &tab[0]);
}
static jac G_TABLE[15];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void init_g_table(void) { make_table(G_TABLE, &GEN); }

/* a*G + b*P, interleaved fixed 4-bit windows (Strauss-Shamir) */
static jac ecmult2(const u256* a, const u256* b, const aff* p) {
  pthread_once(&g_once, init_g_table);
  jac ptab[15];
  jac acc; memset(&acc, 0, sizeof acc); acc.inf = 1;
  int have_p = p != NULL && !p->inf && !is_zero256(b);
  if (have_p) make_table(ptab, p);
  for (int w = 63; w >= 0; w--) {
    for (int k = 0; k < 4; k++) acc = jac_double(&acc);
    unsigned da = (unsigned)(a->l[w / 16] >> (4 * (w % 16))) & 15;
    if (da) acc = jac_add(&acc, &G_TABLE[da - 1]);
    if (have_p) {
      unsigned db = (unsigned)(b->l[w / 16] >> (4 * (w % 16))) & 15;
      if (db) acc = jac_add(&acc, &ptab[db - 1]);
    }
  }
  return acc;
}

/* ------------------------------------------------------------------ public oracle API */

/* SEC 1 v2 §4.1.6 (j = 0 only).  Returns 1 and writes X||Y (64 bytes BE) on success. */
int oracle_ecrecover_pubkey(const uint8_t digest[32], const uint8_t r_be[32], const uint8_t s_be[32], uint8_t v,
                            uint8_t pub_out[64]) {
  u256 r = from_be(r_be), s = from_be(s_be), z = from_be(digest);
  if (v > 1) return 0;
  if (is_zero256(&r) || is_zero256(&s) || cmp256(&r, &FN) >= 0 || cmp256(&s, &FN) >= 0) return 0;
  /* x = r (< n < p, so always a field element) */
  u256 x2 = fp_sqr(&r), x3 = fp_mul(&x2, &r), seven = {{7, 0, 0, 0}}, y2 = fp_add(&x3, &seven), y;
  if (!fp_sqrt(&y, &y2)) return 0;
  if ((y.l[0] & 1) != v) y = fp_neg(&y);
  aff R = {r, y, 0};
  while (cmp256(&z, &FN) >= 0) sub256(&z, &z, &FN);
  u256 rinv = fn_inv(&r);
  u256 u1 = fn_mul(&z, &rinv); u1 = fn_neg(&u1);
  u256 u2 = fn_mul(&s, &rinv);
  jac Q = ecmult2(&u1, &u2, &R);
  aff Qa = jac_to_aff(&Q);
  if (Qa.inf) return 0;
  to_be(&Qa.x, pub_out);
  to_be(&Qa.y, pub_out + 32);
  return 1;
}

int oracle_ecrecover_address(const uint8_t digest[32], const uint8_t r_be[32], const uint8_t s_be[32], uint8_t v,
                             uint8_t addr_out[20]) {
  uint8_t pub[64], h[32];
  if (!oracle_ecrecover_pubkey(digest, r_be, s_be, v, pub)) return 0;
  oracle_keccak256(pub, 64, h);
  memcpy(addr_out, h + 12, 20);
  return 1;
}

/* k*G -> X||Y.  Returns 0 if k = 0 mod n. */
int oracle_pubkey_from_scalar(const uint8_t k_be[32], uint8_t pub_out[64]) {
  u256 k = from_be(k_be), zero = {{0, 0, 0, 0}};
  while (cmp256(&k, &FN) >= 0) sub256(&k, &k, &FN);
  jac Q = ecmult2(&k, &zero, NULL);
  aff Qa = jac_to_aff(&Q);
  if (Qa.inf) return 0;
  to_be(&Qa.x, pub_out);
  to_be(&Qa.y, pub_out + 32);
  return 1;
}

/* a*G + b*P (P given as X||Y, may be NULL) -> X||Y; returns 0 for infinity.  Debug/parity helper. */
int oracle_ecmult2(const uint8_t a_be[32], const uint8_t b_be[32], const uint8_t p_xy[64], uint8_t out[64]) {
  u256 a = from_be(a_be), b = from_be(b_be);
  while (cmp256(&a, &FN) >= 0) sub256(&a, &a, &FN);
  while (cmp256(&b, &FN) >= 0) sub256(&b, &b, &FN);
  aff P; memset(&P, 0, sizeof P); P.inf = 1;
  if (p_xy) { P.x = from_be(p_xy); P.y = from_be(p_xy + 32); P.inf = 0; }
  jac Q = ecmult2(&a, &b, p_xy ? &P : NULL);
  aff Qa = jac_to_aff(&Q);
  if (Qa.inf) return 0;
  to_be(&Qa.x, out);
  to_be(&Qa.y, out + 32);
  return 1;
}

/* ECDSA sign with a caller-supplied nonce (RFC 6979 nonce computed by the Python side).
 * out65 = R||S||V; low_s != 0 normalises s <= n/2.  Returns 0 if the nonce is unusable. */
int oracle_sign_with_k(const uint8_t d_be[32], const uint8_t digest[32], const uint8_t k_be[32], int low_s,
                       uint8_t out65[65]) {
  u256 d = from_be(d_be), z = from_be(digest), k = from_be(k_be), zero = {{0, 0, 0, 0}};
  if (is_zero256(&k) || cmp256(&k, &FN) >= 0) return 0;
  while (cmp256(&z, &FN) >= 0) sub256(&z, &z, &FN);
  jac Rj = ecmult2(&k, &zero, NULL);
  aff R = jac_to_aff(&Rj);
  if (R.inf || cmp256(&R.x, &FN) >= 0) return 0;
  u256 r = R.x;
  if (is_zero256(&r)) return 0;
  u256 kinv = fn_inv(&k), rd = fn_mul(&r, &d), sum;
  if (add256(&sum, &z, &rd) || cmp256(&sum, &FN) >= 0) sub256(&sum, &sum, &FN);
  u256 s = fn_mul(&kinv, &sum);
  if (is_zero256(&s)) return 0;
  uint8_t v = (uint8_t)(R.y.l[0] & 1);
  if (low_s) {
    u256 half = {{0xDFE92F46681B20A0ULL, 0x5D576E7357A4501DULL, 0xFFFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL}};
    if (cmp256(&s, &half) > 0) { s = fn_neg(&s); v ^= 1; }
  }
  to_be(&r, out65);
  to_be(&s, out65 + 32);
  out65[64] = v;
  return 1;
}

/* Field / scalar helpers exposed so the CUDA primitives can be parity-tested one by one. */
void oracle_fp_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  u256 x = from_be(a), y = from_be(b);
  while (cmp256(&x, &FP) >= 0) sub256(&x, &x, &FP);
  while (cmp256(&y, &FP) >= 0) sub256(&y, &y, &FP);
  u256 r = fp_mul(&x, &y);
  to_be(&r, out);
}
void oracle_fp_inv(const uint8_t a[32], uint8_t out[32]) {
  u256 x = from_be(a);
  while (cmp256(&x, &FP) >= 0) sub256(&x, &x, &FP);
  u256 r = fp_inv(&x);
  to_be(&r, out);
}
int oracle_fp_sqrt(const uint8_t a[32], uint8_t out[32]) {
  u256 x = from_be(a), r;
  while (cmp256(&x, &FP) >= 0) sub256(&x, &x, &FP);
  int ok = fp_sqrt(&r, &x);
  to_be(&r, out);
  return ok;
}
void oracle_fn_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  u256 x = from_be(a), y = from_be(b);
  u256 r = fn_mul(&x, &y);
  to_be(&r, out);
}
void oracle_fn_inv(const uint8_t a[32], uint8_t out[32]) {
  u256 x = from_be(a);
  while (cmp256(&x, &FN) >= 0) sub256(&x, &x, &FN);
  u256 r = fn_inv(&x);
  to_be(&r, out);
}

/* ---- batch verification over the product's own packed item records (include/ibft_verify.h).
 * Verdict for one item restates the embedder contract:
 *   IsValidValidator (core/backend.go:41-45): signer recovered from msg.Signature over
 *     Keccak-256(PayloadNoSig) equals msg.From, and From is in the validator set of the height;
 *   IsValidCommittedSeal (core/backend.go:53-55): signer recovered from seal.Signature over
 *     Keccak-256(proposalHash || 0x02) equals seal.Signer, and Signer is in the validator set. */
static int cmp_addr(const void* a, const void* b) { return memcmp(a, b, 20); }
/* `sorted` != 0: the table is sorted (oracle_verify_batch sorts private copies once per call) */
static int addr_in_table(const uint8_t* table, uint32_t n, const uint8_t addr[20], int sorted) {
  if (sorted) return bsearch(addr, table, n, 20, cmp_addr) != NULL ? 0 : -1;
  for (uint32_t i = 0; i < n; i++)
    if (memcmp(table + 20 * (size_t)i, addr, 20) == 0) return (int)i;
  return -1;
}

int oracle_item_digest(const ibft_sig_item* it, const uint8_t* arena, size_t arena_len, uint8_t z[32]) {
  switch (it->kind) {
    case IBFT_KIND_DIGEST: memcpy(z, it->digest, 32); return 1;
    case IBFT_KIND_PAYLOAD:
      if ((size_t)it->payload_off + it->payload_len > arena_len) return 0;
      oracle_keccak256(arena + it->payload_off, it->payload_len, z);
      return 1;
    case IBFT_KIND_PAYLOAD2: { /* PayloadNoSig as two spans (a ROUND_CHANGE head + its shared prepared certificate) */
      uint64_t off2 = 0; uint32_t len2 = 0;
      for (int i = 0; i < 8; i++) off2 |= (uint64_t)it->digest[i] << (8 * i);
      for (int i = 0; i < 4; i++) len2 |= (uint32_t)it->digest[8 + i] << (8 * i);
      if ((size_t)it->payload_off + it->payload_len > arena_len || off2 > arena_len || (size_t)len2 > arena_len - off2) return 0;
      size_t total = (size_t)it->payload_len + len2;
      uint8_t* buf = (uint8_t*)malloc(total ? total : 1);
      if (!buf) return 0;
      memcpy(buf, arena + it->payload_off, it->payload_len);
      memcpy(buf + it->payload_len, arena + off2, len2);
      oracle_keccak256(buf, total, z);
      free(buf);
      return 1;
    }
    case IBFT_KIND_SEAL: {
      uint8_t buf[33];
      memcpy(buf, it->digest, 32);
      buf[32] = 0x02; /* proto.MessageType_COMMIT, messages/proto/messages.proto:10 */
      oracle_keccak256(buf, 33, z);
      return 1;
    }
    default: return 0;
  }
}

static int verify_item(const ibft_sig_item* it, const uint8_t* arena, size_t arena_len, const uint8_t* table,
                       uint32_t table_n, int sorted, uint8_t recovered[20]) {
  uint8_t z[32], addr[20];
  if (recovered) memset(recovered, 0, 20);
  if (!oracle_item_digest(it, arena, arena_len, z)) return 0;
  if (!oracle_ecrecover_address(z, it->r, it->s, it->v, addr)) return 0;
  if (recovered) memcpy(recovered, addr, 20);
  if (memcmp(addr, it->signer, 20) != 0) return 0;
  if (table && addr_in_table(table, table_n, addr, sorted) < 0) return 0;
  return 1;
}
int oracle_verify_item(const ibft_sig_item* it, const uint8_t* arena, size_t arena_len, const uint8_t* table,
                       uint32_t table_n, uint8_t recovered[20]) {
  return verify_item(it, arena, arena_len, table, table_n, 0, recovered);
}

typedef struct {
  const ibft_sig_item* items; uint32_t lo, hi;
  const uint8_t* arena; size_t arena_len;
  const uint8_t* const* tables; const uint32_t* table_n; const uint16_t* group_table; uint32_t n_groups;
  uint8_t* verdict; /* one byte per item (threads never share a word) */
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (uint32_t i = j->lo; i < j->hi; i++) {
    const ibft_sig_item* it = &j->items[i];
    const uint8_t* tab = NULL; uint32_t tn = 0;
    /* an item whose group index is outside the call's groups belongs to no quorum domain and no validator set: verdict 0
     * (IsValidValidator needs "one of the validators at the height in message", core/backend.go:41-45) */
    if (j->group_table && it->group >= j->n_groups) { j->verdict[i] = 0; continue; }
    if (j->group_table && j->group_table[it->group] != 0xFFFF) {
      tab = j->tables[j->group_table[it->group]];
      tn = j->table_n[j->group_table[it->group]];
    }
    j->verdict[i] = (uint8_t)verify_item(it, j->arena, j->arena_len, tab, tn, 1, NULL);
  }
  return NULL;
}

/* Multi-threaded batch verify: the "goroutine-parallel CPU verify" baseline the north star asks
 * for (one worker per host core, contiguous slices).  bitmap gets bit i%32 of word i/32. */
int oracle_verify_batch(const ibft_sig_item* items, uint32_t n, const uint8_t* arena, size_t arena_len,
                        const uint8_t* const* tables, const uint32_t* table_n, const uint16_t* group_table,
                        uint32_t n_groups, int n_threads, uint32_t* bitmap) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_once(&g_once, init_g_table);
  uint8_t* verdict = (uint8_t*)calloc(n ? n : 1, 1);
  if (!verdict) return -1;
  /* sorted private copies of the validator tables (set membership by binary search) */
  uint32_t n_tables = 0;
  for (uint32_t g = 0; group_table && g < n_groups; g++)
    if (group_table[g] != 0xFFFF && group_table[g] + 1u > n_tables) n_tables = group_table[g] + 1u;
  uint8_t** sorted_tabs = (uint8_t**)calloc(n_tables ? n_tables : 1, sizeof(uint8_t*));
  for (uint32_t t = 0; t < n_tables; t++) {
    sorted_tabs[t] = (uint8_t*)malloc((size_t)table_n[t] * 20 + 1);
    memcpy(sorted_tabs[t], tables[t], (size_t)table_n[t] * 20);
    qsort(sorted_tabs[t], table_n[t], 20, cmp_addr);
  }
  tables = (const uint8_t* const*)sorted_tabs;
  pthread_t th[256];
  job_t jobs[256];
  uint32_t per = (n + (uint32_t)n_threads - 1) / (uint32_t)n_threads;
  int started = 0;
  for (int t = 0; t < n_threads; t++) {
    uint32_t lo = (uint32_t)t * per, hi = lo + per > n ? n : lo + per;
    if (lo >= hi) break;
    jobs[t] = (job_t){items, lo, hi, arena, arena_len, tables, table_n, group_table, n_groups, verdict};
    if (n_threads == 1) worker(&jobs[t]);
    else pthread_create(&th[t], NULL, worker, &jobs[t]);
    started++;
  }
  if (n_threads > 1)
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
  memset(bitmap, 0, ((size_t)n + 31) / 32 * 4);
  for (uint32_t i = 0; i < n; i++)
    if (verdict[i]) bitmap[i / 32] |= 1u << (i % 32);
  free(verdict);
  for (uint32_t t = 0; t < n_tables; t++) free(sorted_tabs[t]);
  free(sorted_tabs);
  return 0;
}

/* ---- bulk workload generation (tests/workloads.py: the full-size BASELINE configs 4 and 5 need 10^5 keys and signatures;
 * one ctypes call per key from Python would take minutes).  Same primitives as above, spread over threads. ---- */
typedef void (*range_fn)(uint32_t lo, uint32_t hi, void* ctx);
typedef struct { range_fn fn; uint32_t lo, hi; void* ctx; } range_job;
static void* range_worker(void* a) { range_job* j = (range_job*)a; j->fn(j->lo, j->hi, j->ctx); return NULL; }
static void parallel_ranges(uint32_t n, int n_threads, range_fn fn, void* ctx) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_once(&g_once, init_g_table);
  pthread_t th[256];
  range_job jobs[256];
  uint32_t per = (n + (uint32_t)n_threads - 1) / (uint32_t)n_threads;
  int started = 0;
  for (int t = 0; t < n_threads; t++) {
    uint32_t lo = (uint32_t)t * per, hi = lo + per > n ? n : lo + per;
    if (lo >= hi) break;
    jobs[t] = (range_job){fn, lo, hi, ctx};
    if (n_threads == 1) range_worker(&jobs[t]);
    else pthread_create(&th[t], NULL, range_worker, &jobs[t]);
    started++;
  }
  if (n_threads > 1)
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

/* privkey_i = Keccak-256("ibft-b200-validator" || u32_be(seed) || u32_be(i)) mod (n-1) + 1   (SURVEY.md §8d), i = first.. */
void oracle_privkeys(uint32_t seed, uint32_t first, uint32_t n, uint8_t* out32) {
  u256 one = {{1, 0, 0, 0}}, nm1;
  sub256(&nm1, &FN, &one);
  for (uint32_t k = 0; k < n; k++) {
    uint32_t i = first + k;
    uint8_t buf[27] = "ibft-b200-validator";
    buf[19] = (uint8_t)(seed >> 24); buf[20] = (uint8_t)(seed >> 16); buf[21] = (uint8_t)(seed >> 8); buf[22] = (uint8_t)seed;
    buf[23] = (uint8_t)(i >> 24); buf[24] = (uint8_t)(i >> 16); buf[25] = (uint8_t)(i >> 8); buf[26] = (uint8_t)i;
    uint8_t h[32];
    oracle_keccak256(buf, 27, h);
    u256 x = from_be(h);
    while (cmp256(&x, &nm1) >= 0) sub256(&x, &x, &nm1);
    add256(&x, &x, &one);
    to_be(&x, out32 + 32 * (size_t)k);
  }
}

typedef struct { const uint8_t* privs; uint8_t* out20; } addr_ctx;
static void addr_range(uint32_t lo, uint32_t hi, void* c) {
  addr_ctx* a = (addr_ctx*)c;
  for (uint32_t i = lo; i < hi; i++) {
    uint8_t pub[64], h[32];
    if (!oracle_pubkey_from_scalar(a->privs + 32 * (size_t)i, pub)) { memset(a->out20 + 20 * (size_t)i, 0, 20); continue; }
    oracle_keccak256(pub, 64, h);
    memcpy(a->out20 + 20 * (size_t)i, h + 12, 20);
  }
}
/* address_i = Keccak-256(X||Y)[12:] of privs[i]*G */
void oracle_addresses(const uint8_t* privs32, uint32_t n, uint8_t* out20, int n_threads) {
  addr_ctx c = {privs32, out20};
  parallel_ranges(n, n_threads, addr_range, &c);
}

typedef struct { const uint8_t* privs; const uint8_t* digests; uint8_t* out65; } sign_ctx;
static void sign_range(uint32_t lo, uint32_t hi, void* c) {
  sign_ctx* s = (sign_ctx*)c;
  for (uint32_t i = lo; i < hi; i++) {
    const uint8_t* d = s->privs + 32 * (size_t)i;
    const uint8_t* z = s->digests + 32 * (size_t)i;
    uint8_t* o = s->out65 + 65 * (size_t)i;
    int ok = 0;
    /* derived nonce, the rule of the engine's k_sign: k = Keccak-256(d || z || ctr), ctr = 0, 1, ... until usable */
    for (uint32_t ctr = 0; ctr < 4 && !ok; ctr++) {
      uint8_t buf[65], k[32];
      memcpy(buf, d, 32); memcpy(buf + 32, z, 32); buf[64] = (uint8_t)ctr;
      oracle_keccak256(buf, 65, k);
      ok = oracle_sign_with_k(d, z, k, 1, o);
    }
    if (!ok) memset(o, 0, 65);
  }
}
/* sigs[i] = ECDSA(privs[i], digests[i]) with the deterministic Keccak-derived nonce, low-s, R||S||V */
void oracle_sign_derived_batch(const uint8_t* privs32, const uint8_t* digests32, uint32_t n, uint8_t* out65, int n_threads) {
  sign_ctx c = {privs32, digests32, out65};
  parallel_ranges(n, n_threads, sign_range, &c);
}

typedef struct { const uint8_t* arena; const uint64_t* offs; const uint32_t* lens; uint8_t* out32; } kb_ctx;
static void kb_range(uint32_t lo, uint32_t hi, void* c) {
  kb_ctx* k = (kb_ctx*)c;
  for (uint32_t i = lo; i < hi; i++) oracle_keccak256(k->arena + k->offs[i], k->lens[i], k->out32 + 32 * (size_t)i);
}
void oracle_keccak256_batch(const uint8_t* arena, const uint64_t* offs, const uint32_t* lens, uint32_t n, uint8_t* out32, int n_threads) {
  kb_ctx c = {arena, offs, lens, out32};
  parallel_ranges(n, n_threads, kb_range, &c);
}
