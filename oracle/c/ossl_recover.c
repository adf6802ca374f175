/* ossl_recover.c -- a second, TUNED-LIBRARY CPU arm for bench.py's cpu_baseline: the same recover + compare + membership
 * verdict as oracle_verify_batch (ibft_oracle.c), with the point arithmetic done by OpenSSL 3 (EC_POINT_set_compressed_coordinates
 * for the lift, EC_POINT_mul for u1*G + u2*R) and this repo's Keccak.  BASELINE.md §3 planned this arm; it is test / bench
 * infrastructure only (oracle/__init__.py) and is cross-checked bit for bit against the plain port in tests/test_oracle_crypto.py.
 * Conventions identical to the port: x = r only, v in {0,1}, 1 <= r,s < n, high-s accepted, address = Keccak-256(X||Y)[12:]. */
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/obj_mac.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ibft_verify.h"

void oracle_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);
int oracle_item_digest(const ibft_sig_item* it, const uint8_t* arena, size_t arena_len, uint8_t z[32]);

typedef struct {
  EC_GROUP* g;
  BN_CTX* ctx;
  BIGNUM *n, *r, *s, *z, *rinv, *u1, *u2, *x, *y;
  EC_POINT *R, *Q;
} ossl_state;

static int st_init(ossl_state* t) {
  memset(t, 0, sizeof *t);
  t->g = EC_GROUP_new_by_curve_name(NID_secp256k1);
  t->ctx = BN_CTX_new();
  if (!t->g || !t->ctx) return 0;
  t->n = BN_new(); t->r = BN_new(); t->s = BN_new(); t->z = BN_new(); t->rinv = BN_new(); t->u1 = BN_new(); t->u2 = BN_new();
  t->x = BN_new(); t->y = BN_new();
  t->R = EC_POINT_new(t->g); t->Q = EC_POINT_new(t->g);
  return EC_GROUP_get_order(t->g, t->n, t->ctx) == 1;
}
static void st_free(ossl_state* t) {
  BN_free(t->n); BN_free(t->r); BN_free(t->s); BN_free(t->z); BN_free(t->rinv); BN_free(t->u1); BN_free(t->u2); BN_free(t->x); BN_free(t->y);
  EC_POINT_free(t->R); EC_POINT_free(t->Q);
  BN_CTX_free(t->ctx); EC_GROUP_free(t->g);
}

static int ossl_recover_address(ossl_state* t, const uint8_t z32[32], const uint8_t r32[32], const uint8_t s32[32], uint8_t v,
                                uint8_t addr[20]) {
  if (v > 1) return 0;
  BN_bin2bn(r32, 32, t->r); BN_bin2bn(s32, 32, t->s); BN_bin2bn(z32, 32, t->z);
  if (BN_is_zero(t->r) || BN_is_zero(t->s) || BN_cmp(t->r, t->n) >= 0 || BN_cmp(t->s, t->n) >= 0) return 0;
  if (EC_POINT_set_compressed_coordinates(t->g, t->R, t->r, v, t->ctx) != 1) return 0; /* r is not an abscissa */
  if (!BN_mod_inverse(t->rinv, t->r, t->n, t->ctx)) return 0;
  BN_nnmod(t->z, t->z, t->n, t->ctx);
  BN_mod_mul(t->u1, t->z, t->rinv, t->n, t->ctx);
  BN_sub(t->u1, t->n, t->u1);                      /* u1 = -z/r */
  BN_nnmod(t->u1, t->u1, t->n, t->ctx);
  BN_mod_mul(t->u2, t->s, t->rinv, t->n, t->ctx);  /* u2 = s/r */
  if (EC_POINT_mul(t->g, t->Q, t->u1, t->R, t->u2, t->ctx) != 1) return 0;
  if (EC_POINT_is_at_infinity(t->g, t->Q)) return 0;
  if (EC_POINT_get_affine_coordinates(t->g, t->Q, t->x, t->y, t->ctx) != 1) return 0;
  uint8_t xy[64], h[32];
  BN_bn2binpad(t->x, xy, 32);
  BN_bn2binpad(t->y, xy + 32, 32);
  oracle_keccak256(xy, 64, h);
  memcpy(addr, h + 12, 20);
  return 1;
}

static int cmp20(const void* a, const void* b) { return memcmp(a, b, 20); }

typedef struct {
  const ibft_sig_item* items; uint32_t lo, hi; const uint8_t* arena; size_t arena_len;
  const uint8_t* table; uint32_t table_n; uint8_t* verdict;
} ojob;
static void* oworker(void* a) {
  ojob* j = (ojob*)a;
  ossl_state t;
  if (!st_init(&t)) return NULL;
  for (uint32_t i = j->lo; i < j->hi; i++) {
    const ibft_sig_item* it = &j->items[i];
    uint8_t z[32], addr[20];
    int ok = oracle_item_digest(it, j->arena, j->arena_len, z) && ossl_recover_address(&t, z, it->r, it->s, it->v, addr) &&
             memcmp(addr, it->signer, 20) == 0;
    if (ok && j->table) ok = bsearch(addr, j->table, j->table_n, 20, cmp20) != NULL;
    j->verdict[i] = (uint8_t)ok;
  }
  st_free(&t);
  return NULL;
}

/* one validator table (may be NULL) for all items; bitmap gets bit i%32 of word i/32 */
int ossl_verify_batch(const ibft_sig_item* items, uint32_t n, const uint8_t* arena, size_t arena_len, const uint8_t* table,
                      uint32_t table_n, int n_threads, uint32_t* bitmap) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  uint8_t* verdict = (uint8_t*)calloc(n ? n : 1, 1);
  uint8_t* sorted = NULL;
  if (table) {
    sorted = (uint8_t*)malloc((size_t)table_n * 20 + 1);
    memcpy(sorted, table, (size_t)table_n * 20);
    qsort(sorted, table_n, 20, cmp20);
  }
  pthread_t th[256];
  ojob jobs[256];
  uint32_t per = (n + (uint32_t)n_threads - 1) / (uint32_t)n_threads;
  int started = 0;
  for (int k = 0; k < n_threads; k++) {
    uint32_t lo = (uint32_t)k * per, hi = lo + per > n ? n : lo + per;
    if (lo >= hi) break;
    jobs[k] = (ojob){items, lo, hi, arena, arena_len, sorted, table_n, verdict};
    if (n_threads == 1) oworker(&jobs[k]);
    else pthread_create(&th[k], NULL, oworker, &jobs[k]);
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
