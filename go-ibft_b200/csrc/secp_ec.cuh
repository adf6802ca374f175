// secp_ec.cuh -- secp256k1 group law (Jacobian coordinates, a = 0) and the double-scalar multiplication
// u1*G + u2*R that dominates public-key recovery.
//
// SIMT-first formulation (this is the part a CPU library would do differently):
//   * GLV: both scalars are split into two ~128-bit halves (secp_scalar.cuh), so one chain of 132 doublings
//     serves four digit streams;
//   * signed FIXED windows (Booth digits) instead of wNAF: every lane of a warp performs its additions at the
//     same loop positions, so a warp never executes an addition for a minority of its lanes (wNAF's
//     data-dependent digit positions would leave ~5/6 of the lanes idle in every addition);
//   * the generator table ({1..2^(WG-1)} * G, affine, with beta*x alongside) is staged in shared memory once per
//     CTA; the per-signature table {1..8} * R lives in thread-local memory;
//   * exceptional cases (P + P, P - P, infinity) are handled exactly -- adversarial signatures can reach them
//     and the verdict must be bit-exact with the oracle -- but by rarely-taken branches.
#pragma once
#include "secp_fe.cuh"
#include "secp_scalar.cuh"

namespace ibft {

struct jac {
  fe x, y, z;
  bool inf;
};
struct aff {
  fe x, y;
};

// beta: cube root of unity mod p with lambda*(x,y) = (beta*x, y)
IBFT_HD fe fe_beta() {
  fe b;
  const uint32_t v[8] = {0x719501EEu, 0xC1396C28u, 0x12F58995u, 0x9CF04975u, 0xAC3434E9u, 0x6E64479Eu, 0x657C0710u, 0x7AE96A2Bu};
#pragma unroll
  for (int i = 0; i < 8; i++) b.v[i] = v[i];
  return b;
}

// dbl-2009-l: 2M + 5S
IBFT_HD jac jac_double(const jac& p) {
  jac r;
  // Y = 0 never happens on secp256k1 (no points of order 2), so no exceptional case besides infinity.
  fe a = fe_sqr(p.x);
  fe b = fe_sqr(p.y);
  fe c = fe_sqr(b);
  fe t = fe_add(p.x, b);
  t = fe_sqr(t);
  t = fe_sub(t, a);
  t = fe_sub(t, c);
  fe d = fe_dbl(t);
  fe e = fe_add(fe_dbl(a), a);
  fe f = fe_sqr(e);
  r.x = fe_sub(f, fe_dbl(d));
  fe c8 = fe_dbl(fe_dbl(fe_dbl(c)));
  r.y = fe_sub(fe_mul(e, fe_sub(d, r.x)), c8);
  r.z = fe_dbl(fe_mul(p.y, p.z));
  r.inf = p.inf;
  return r;
}

// p + (qx, qy) with q affine (never infinity): 8M + 3S
IBFT_HD jac jac_add_affine(const jac& p, const fe& qx, const fe& qy) {
  jac r;
  if (p.inf) {
    r.x = qx;
    r.y = qy;
    r.z = fe_from_u32(1);
    r.inf = false;
    return r;
  }
  fe z1z1 = fe_sqr(p.z);
  fe u2 = fe_mul(qx, z1z1);
  fe s2 = fe_mul(fe_mul(qy, p.z), z1z1);
  fe h = fe_sub(u2, p.x);
  fe rr = fe_sub(s2, p.y);
  if (fe_is_zero(h)) {
    if (fe_is_zero(rr)) return jac_double(p);  // P + P
    r = p;
    r.inf = true;  // P + (-P)
    return r;
  }
  fe hh = fe_sqr(h);
  fe hhh = fe_mul(h, hh);
  fe v = fe_mul(p.x, hh);
  r.x = fe_sub(fe_sub(fe_sqr(rr), hhh), fe_dbl(v));
  r.y = fe_sub(fe_mul(rr, fe_sub(v, r.x)), fe_mul(p.y, hhh));
  r.z = fe_mul(p.z, h);
  r.inf = false;
  return r;
}

// p + q, both Jacobian (q never infinity): 12M + 4S
IBFT_HD jac jac_add(const jac& p, const fe& qx, const fe& qy, const fe& qz) {
  jac r;
  if (p.inf) {
    r.x = qx;
    r.y = qy;
    r.z = qz;
    r.inf = false;
    return r;
  }
  fe z1z1 = fe_sqr(p.z);
  fe z2z2 = fe_sqr(qz);
  fe u1 = fe_mul(p.x, z2z2);
  fe u2 = fe_mul(qx, z1z1);
  fe s1 = fe_mul(fe_mul(p.y, qz), z2z2);
  fe s2 = fe_mul(fe_mul(qy, p.z), z1z1);
  fe h = fe_sub(u2, u1);
  fe rr = fe_sub(s2, s1);
  if (fe_is_zero(h)) {
    if (fe_is_zero(rr)) return jac_double(p);
    r = p;
    r.inf = true;
    return r;
  }
  fe hh = fe_sqr(h);
  fe hhh = fe_mul(h, hh);
  fe v = fe_mul(u1, hh);
  r.x = fe_sub(fe_sub(fe_sqr(rr), hhh), fe_dbl(v));
  r.y = fe_sub(fe_mul(rr, fe_sub(v, r.x)), fe_mul(s1, hhh));
  r.z = fe_mul(fe_mul(p.z, qz), h);
  r.inf = false;
  return r;
}

// ------------------------------------------------------------------------------------------------
// u1*G + u2*R
// ------------------------------------------------------------------------------------------------
#ifndef IBFT_WG
#define IBFT_WG 8  // generator window: 2^(WG-1) table entries of (x, y, beta*x)
#endif
#define IBFT_WR 4  // per-signature window: table {1..8} * R
#define IBFT_GTAB_ENTRIES (1 << (IBFT_WG - 1))
#define IBFT_NWIN_R 33  // ceil(130 / 4)
static_assert(IBFT_WG % IBFT_WR == 0, "generator window must be a multiple of the R window");

// Generator table accessor: entry i (0-based) = (i+1)*G as 24 words x[8] y[8] bx[8].
struct gtab_view {
  const uint32_t* base;  // shared or global memory
  IBFT_HD void load(int idx, bool lambda, fe& x, fe& y) const {
    const uint32_t* e = base + 24 * idx;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      x.v[i] = lambda ? e[16 + i] : e[i];
      y.v[i] = e[8 + i];
    }
  }
};

struct rtab_entry {
  fe x, y, z, bx;
};

// Returns u1*G + u2*R (R affine, on the curve) as a Jacobian point.  u1, u2 in [0, n).
IBFT_HD jac ecmult_double(const sc& u1, const sc& u2, const aff& R, const gtab_view& G) {
  glv_half g1, g2, r1, r2;
  glv_split(u1, g1, g2);
  glv_split(u2, r1, r2);
  uint32_t kg1[6], kg2[6], kr1[6], kr2[6];
#pragma unroll
  for (int i = 0; i < 5; i++) { kg1[i] = g1.k[i]; kg2[i] = g2.k[i]; kr1[i] = r1.k[i]; kr2[i] = r2.k[i]; }
  kg1[5] = kg2[5] = kr1[5] = kr2[5] = 0;

  // {1..8} * R
  rtab_entry tab[8];
  {
    fe beta = fe_beta();
    jac p1;
    p1.x = R.x; p1.y = R.y; p1.z = fe_from_u32(1); p1.inf = false;
    jac p2 = jac_double(p1);
    jac p3 = jac_add_affine(p2, R.x, R.y);
    jac p4 = jac_double(p2);
    jac p5 = jac_add_affine(p4, R.x, R.y);
    jac p6 = jac_double(p3);
    jac p7 = jac_add_affine(p6, R.x, R.y);
    jac p8 = jac_double(p4);
#define IBFT_SETTAB(i, p)                                 \
  tab[i].x = (p).x; tab[i].y = (p).y; tab[i].z = (p).z;   \
  tab[i].bx = fe_mul((p).x, beta);
    IBFT_SETTAB(0, p1) IBFT_SETTAB(1, p2) IBFT_SETTAB(2, p3) IBFT_SETTAB(3, p4)
    IBFT_SETTAB(4, p5) IBFT_SETTAB(5, p6) IBFT_SETTAB(6, p7) IBFT_SETTAB(7, p8)
#undef IBFT_SETTAB
  }

  jac acc;
  acc.x = fe_zero(); acc.y = fe_zero(); acc.z = fe_zero();
  acc.inf = true;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int j = IBFT_NWIN_R - 1; j >= 0; j--) {
    if (!acc.inf) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int t = 0; t < IBFT_WR; t++) acc = jac_double(acc);
    }
    // R streams (every window)
    {
      int d = booth_digit<IBFT_WR>(kr1, j);
      if (d != 0) {
        int idx = (d < 0 ? -d : d) - 1;
        bool neg = (d < 0) != r1.neg;
        fe y = tab[idx].y;
        if (neg) y = fe_neg(y);
        acc = jac_add(acc, tab[idx].x, y, tab[idx].z);
      }
    }
    {
      int d = booth_digit<IBFT_WR>(kr2, j);
      if (d != 0) {
        int idx = (d < 0 ? -d : d) - 1;
        bool neg = (d < 0) != r2.neg;
        fe y = tab[idx].y;
        if (neg) y = fe_neg(y);
        acc = jac_add(acc, tab[idx].bx, y, tab[idx].z);
      }
    }
    // G streams (every WG/WR-th window)
    if (j % (IBFT_WG / IBFT_WR) == 0) {
      int jg = j / (IBFT_WG / IBFT_WR);
      {
        int d = booth_digit<IBFT_WG>(kg1, jg);
        if (d != 0) {
          fe x, y;
          G.load((d < 0 ? -d : d) - 1, false, x, y);
          if ((d < 0) != g1.neg) y = fe_neg(y);
          acc = jac_add_affine(acc, x, y);
        }
      }
      {
        int d = booth_digit<IBFT_WG>(kg2, jg);
        if (d != 0) {
          fe x, y;
          G.load((d < 0 ? -d : d) - 1, true, x, y);
          if ((d < 0) != g2.neg) y = fe_neg(y);
          acc = jac_add_affine(acc, x, y);
        }
      }
    }
  }
  return acc;
}

}  // namespace ibft
