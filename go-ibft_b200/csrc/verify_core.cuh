// verify_core.cuh -- one signature check: (r, s, v, z) -> recovered signer address.
//
// This is the arithmetic behind core.Verifier.IsValidValidator (reference core/backend.go:41-45; callers
// core/ibft.go:735,1128,1213,1220) and core.Verifier.IsValidCommittedSeal (core/backend.go:53-55; caller
// core/ibft.go:943).  Conventions are those of SURVEY.md §8(c) and oracle/secp256k1.py: SEC 1 v2 §4.1.6 recovery
// with x = r only, 1 <= r,s < n, v in {0,1}, high-s accepted, address = Keccak-256(X||Y)[12:].
#pragma once
#include "../../include/ibft_verify.h"
#include "keccak.cuh"
#include "secp_ec.cuh"
#include "secp_modinv.cuh"

namespace ibft {

// modular inversions: safegcd divsteps by default, the Fermat ladders with -DIBFT_FERMAT (kept as a cross-check)
#if defined(IBFT_FERMAT)
#define IBFT_FE_INV fe_inv_fermat
#define IBFT_SC_INV sc_inv_fermat
#else
#define IBFT_FE_INV fe_inv_safegcd
#define IBFT_SC_INV sc_inv_safegcd
#endif

IBFT_HD fe fe_inv_for_table(const fe& a) { return IBFT_FE_INV(a); }

// digest of one item (kinds of include/ibft_verify.h).  Returns false for an unknown kind / out-of-range payload.
IBFT_HD bool item_digest(const ibft_sig_item& it, const uint8_t* arena, size_t arena_len, uint8_t* z) {
  switch (it.kind) {
    case IBFT_KIND_DIGEST:
#pragma unroll
      for (int i = 0; i < 32; i++) z[i] = it.digest[i];
      return true;
    case IBFT_KIND_PAYLOAD:
      if ((size_t)it.payload_off + it.payload_len > arena_len) return false;
      keccak256_bytes(arena + it.payload_off, it.payload_len, z);
      return true;
    case IBFT_KIND_SEAL: {
      uint8_t buf[33];
#pragma unroll
      for (int i = 0; i < 32; i++) buf[i] = it.digest[i];
      buf[32] = 0x02;  // proto.MessageType_COMMIT (reference messages/proto/messages.proto:10)
      keccak256_bytes(buf, 33, z);
      return true;
    }
    default:
      return false;
  }
}

// Recover the signer of (r, s, v) over digest z.  Returns false when the signature is invalid; addr20 then zero.
IBFT_HD bool ecrecover_address(const uint8_t* r_be, const uint8_t* s_be, uint8_t v, const uint8_t* z_be,
                               const gtab_view& G, const rtab_view& T, uint8_t* addr20) {
#pragma unroll
  for (int i = 0; i < 20; i++) addr20[i] = 0;
  sc r = sc_from_be(r_be), s = sc_from_be(s_be);
  if (v > 1) return false;
  if (sc_is_zero(r) || sc_is_zero(s) || sc_ge_n(r) || sc_ge_n(s)) return false;
  // R = lift_x(r, v)   (r < n < p: always a canonical field element)
  fe x;
#pragma unroll
  for (int i = 0; i < 8; i++) x.v[i] = r.v[i];
  fe y2 = fe_add(fe_mul(fe_sqr(x), x), fe_from_u32(7));
  fe y = fe_sqrt_candidate(y2);
  if (!fe_equal(fe_sqr(y), y2)) return false;  // x^3 + 7 is a non-residue: r is not an abscissa
  y = fe_normalize(y);
  if ((y.v[0] & 1u) != (uint32_t)v) y = fe_normalize(fe_neg(y));
  aff R;
  R.x = x;
  R.y = y;
  // u1 = -z/r, u2 = s/r  (mod n)
  sc z = sc_reduce_once(sc_from_be(z_be));
  sc rinv = IBFT_SC_INV(r);
  sc u1 = sc_neg(sc_mul(z, rinv));
  sc u2 = sc_mul(s, rinv);
  jac Q = ecmult_double(u1, u2, R, G, T);
  if (Q.inf || fe_is_zero(Q.z)) return false;
  fe zi = IBFT_FE_INV(Q.z);
  fe zi2 = fe_sqr(zi);
  fe qx = fe_normalize(fe_mul(Q.x, zi2));
  fe qy = fe_normalize(fe_mul(Q.y, fe_mul(zi2, zi)));
  keccak256_xy_address(qx, qy, addr20);
  return true;
}

// ECDSA signing with a given nonce (the MessageConstructor side: reference core/backend.go:12-34 requires every built
// message to be signed and BuildCommitMessage to create a committed seal).  sig65 = R||S||V, s normalised to the low half.
// Returns false when the nonce is unusable (k = 0 mod n, R.x >= n, r = 0 or s = 0) -- same rule as oracle_sign_with_k.
IBFT_HD bool ecdsa_sign(const uint8_t* d_be, const uint8_t* z_be, const uint8_t* k_be, const gtab_view& G, const rtab_view& T,
                        uint8_t* sig65) {
#pragma unroll
  for (int i = 0; i < 65; i++) sig65[i] = 0;
  sc k = sc_from_be(k_be);
  if (sc_is_zero(k) || sc_ge_n(k)) return false;
  sc d = sc_reduce_once(sc_from_be(d_be));
  sc z = sc_reduce_once(sc_from_be(z_be));
  aff g1;
  G.load(0, g1.x, g1.y);  // 1*G; its digit streams are all zero (u2 = 0), only the generator streams add
  sc zero;
#pragma unroll
  for (int i = 0; i < 8; i++) zero.v[i] = 0;
  jac Rj = ecmult_double(k, zero, g1, G, T);
  if (Rj.inf || fe_is_zero(Rj.z)) return false;
  fe zi = IBFT_FE_INV(Rj.z);
  fe zi2 = fe_sqr(zi);
  fe rx = fe_normalize(fe_mul(Rj.x, zi2));
  fe ry = fe_normalize(fe_mul(Rj.y, fe_mul(zi2, zi)));
  sc r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = rx.v[i];
  if (sc_ge_n(r) || sc_is_zero(r)) return false;
  sc s = sc_mul(IBFT_SC_INV(k), sc_add(z, sc_mul(r, d)));
  if (sc_is_zero(s)) return false;
  uint8_t v = (uint8_t)(ry.v[0] & 1u);
  if (sc_is_high(s)) {
    s = sc_neg(s);
    v ^= 1;
  }
  sc_to_be(r, sig65);
  sc_to_be(s, sig65 + 32);
  sig65[64] = v;
  return true;
}

}  // namespace ibft
