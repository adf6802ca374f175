// Host emulation of the device arithmetic (TEST ONLY).
//
// Compiles the very same csrc/*.cuh headers with g++ (their portable C++ path) so that the complete
// per-signature pipeline -- GLV split, Booth digits, group law with exceptional cases, Keccak, address derivation --
// can be checked against the oracle on a CPU-only box.  This library is never loaded by the product; the product
// path (csrc/engine.cu) fails loudly without a CUDA device.  The PTX carry-chain multipliers are exercised on the
// GPU by tests/test_gpu_primitives.py through ibft_debug_op.
#include <cstring>

#include "../../go-ibft_b200/csrc/verify_core.cuh"
#include "../../go-ibft_b200/csrc/secp_gtable.inc"

using namespace ibft;

extern "C" {

int emul_verify_item(const ibft_sig_item* it, const uint8_t* arena, size_t arena_len, uint8_t* recovered20) {
  uint8_t addr[20];
  memset(recovered20, 0, 20);
  resolved_item ri;
  bool valid = false;
  int st = resolve_item(*it, arena, arena_len, ri, &valid);
  if (st != IBFT_ITEM_OK) return -1;  // NEEDS_HOST
  if (!valid) return 0;
  gtab_view G{IBFT_GTABLE};
  uint32_t rtab[IBFT_RTAB_WORDS];
  rtab_view T{rtab, 1};
  if (!ecrecover_address(ri.r, ri.s, ri.v, ri.z, G, T, addr)) return 0;
  memcpy(recovered20, addr, 20);
  return memcmp(addr, ri.signer, 20) == 0;
}

void emul_keccak256(const uint8_t* d, uint32_t n, uint8_t* out) { keccak256_bytes(d, n, out); }

// op codes = IBFT_DBG_* of include/ibft_verify.h
int emul_debug_op(int op, const uint8_t* a, const uint8_t* b, const uint8_t* c, uint8_t* out) {
  switch (op) {
    case IBFT_DBG_FE_MUL: fe_to_be(fe_normalize(fe_mul(fe_from_be(a), fe_from_be(b))), out); return 0;
    case IBFT_DBG_FE_SQR: fe_to_be(fe_normalize(fe_sqr(fe_from_be(a))), out); return 0;
    case IBFT_DBG_FE_INV: fe_to_be(fe_normalize(IBFT_FE_INV(fe_from_be(a))), out); return 0;
    case IBFT_DBG_FE_SQRT: fe_to_be(fe_normalize(fe_sqrt_candidate(fe_from_be(a))), out); return 0;
    case IBFT_DBG_FE_ADD: fe_to_be(fe_normalize(fe_add(fe_from_be(a), fe_from_be(b))), out); return 0;
    case IBFT_DBG_FE_SUB: fe_to_be(fe_normalize(fe_sub(fe_from_be(a), fe_from_be(b))), out); return 0;
    case IBFT_DBG_SC_MUL: sc_to_be(sc_mul(sc_from_be(a), sc_from_be(b)), out); return 0;
    case IBFT_DBG_SC_INV: sc_to_be(IBFT_SC_INV(sc_reduce_once(sc_from_be(a))), out); return 0;
    case IBFT_DBG_GLV: {
      glv_half h1, h2;
      glv_split(sc_reduce_once(sc_from_be(a)), h1, h2);
      memset(out, 0, 64);
      for (int i = 0; i < 5; i++)
        for (int j = 0; j < 4; j++) {
          out[4 * i + j] = (uint8_t)(h1.k[i] >> (8 * j));       // little-endian magnitude, 20 bytes
          out[24 + 4 * i + j] = (uint8_t)(h2.k[i] >> (8 * j));
        }
      out[20] = h1.neg;
      out[44] = h2.neg;
      return 0;
    }
    case IBFT_DBG_ECMULT: {  // out = a*G + b*P, P = (c[0..31], c[32..63]) affine on the curve; all-zero out = infinity
      gtab_view G{IBFT_GTABLE};
      aff P;
      P.x = fe_from_be(c);
      P.y = fe_from_be(c + 32);
      uint32_t rtab[IBFT_RTAB_WORDS];
      rtab_view T{rtab, 1};
      jac Q = ecmult_double(sc_reduce_once(sc_from_be(a)), sc_reduce_once(sc_from_be(b)), P, G, T);
      memset(out, 0, 64);
      if (Q.inf || fe_is_zero(Q.z)) return 1;
      fe zi = IBFT_FE_INV(Q.z), zi2 = fe_sqr(zi);
      fe_to_be(fe_normalize(fe_mul(Q.x, zi2)), out);
      fe_to_be(fe_normalize(fe_mul(Q.y, fe_mul(zi2, zi))), out + 32);
      return 0;
    }
    default: return -1;
  }
}
}

extern "C" int emul_sign(const uint8_t* d, const uint8_t* z, const uint8_t* k, uint8_t* sig65) {
  gtab_view G{IBFT_GTABLE};
  uint32_t rtab[IBFT_RTAB_WORDS];
  rtab_view T{rtab, 1};
  return ecdsa_sign(d, z, k, G, T, sig65) ? 1 : 0;
}
