// osb_math.cuh -- per-interface arithmetic of the characteristic LLF flux, shared by every sweep kernel.
//
// Hand-written restatement (NOT a translation of generated code) of
//   opensbli/schemes/spatial/shock_capturing.py:357-536   LLF characteristic pre/post-process
//   opensbli/schemes/spatial/averaging.py:31-114          Simple / Roe interface state
//   opensbli/physical_models/euler_eigensystem.py:57-135  1-D/2-D/3-D eigensystems (identity cosines)
//   opensbli/schemes/spatial/teno.py:57-465               TENO5 / TENO6
//   opensbli/schemes/spatial/weno.py:35-465               WENO5-JS / WENO5-Z
// restructured for the FP64 pipe:
//   * the left-eigenvector products use the structure  L.v = { v0 - S/a^2, -(k x w)/rho, beta(S +- a k.w) }
//     with S = phi^2 v0 - (g-1) u.v_m + (g-1) v_E and w = v_m - u v0, instead of dense 5x5 products;
//   * TENO cut-off without divisions: with D_r = eps + beta_r, N_r = D_r + tau, P_r = N_r prod_{s!=r} D_s,
//       alpha_r / sum(alpha) < C_T   <=>   P_r^6 < C_T sum_s P_s^6     (common factor (prod D)^6 > 0),
//     and the normalisation 1/sum(d_r delta_r) takes one of 2^n-1 values chosen by predicates;
//   * powers by repeated multiplication; halves of the flux splitting folded into the ENO coefficients.
// All functions are __host__ __device__ so the same source is unit-tested on the CPU (tests/csrc).
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define OSB_HD __host__ __device__ __forceinline__
#else
#define OSB_HD inline
#endif

namespace osb {

enum Recon : int { RECON_WENO5_JS = 0, RECON_WENO5_Z = 1, RECON_TENO5 = 2, RECON_TENO6 = 3 };
enum Averaging : int { AVG_SIMPLE = 0, AVG_ROE = 1 };

struct SchemeParams {
  double eps;      // TENO eps (runtime constant `eps`, teno.py:358-365)
  double teno_ct;  // TENO cut-off C_T
  // derived on the host by make_scheme_params():
  double eps16;    // 16*eps: TENO5 works on g = 2f with beta'' = 16 beta(f)
  double kfast5;   // 0.98 (3 C_T)^(-1/6): if 1 + tau/D_min <= kfast every TENO5 stencil passes the cut-off
  double kpass5;   // kfast5 - 1: the same test as  tau <= kpass5 D_r  for every r  (three compares, no minimum)
  unsigned long long *slow_count;   // instrumentation (null in production): number of TENO5 waves that took the full cut-off path
  double kfast6;   // 0.98 (4 C_T)^(-1/6): same for the 4 stencils of TENO6
};

// adaptive TENO (teno.py:430-443): C_T = 10^-e, e = floor(a1 - a2 (1 - (1-theta)^4 (1+4 theta))) from the Ducros sensor.
// e is a small integer, so C_T and the derived constants come from tables built on the host with the same pow().
struct AdaptiveCT {
  int on;
  double a1, a2;
  double ct[16], k5[16], k6[16];
};
OSB_HD int adaptive_exponent(const AdaptiveCT &ad, double theta) {
  const double om = 1.0 - theta;
  const double e = floor(ad.a1 - ad.a2 * (-(om * om) * (om * om) * (4.0 * theta + 1.0) + 1.0));
  const int i = (int)e;
  return i < 0 ? 0 : (i > 15 ? 15 : i);
}
inline AdaptiveCT make_adaptive_ct(bool on, double a1, double a2) {
  AdaptiveCT ad;
  ad.on = on ? 1 : 0; ad.a1 = a1; ad.a2 = a2;
  for (int e = 0; e < 16; e++) {
    ad.ct[e] = pow(10.0, -(double)e);
    ad.k5[e] = 0.98 * pow(3.0 * ad.ct[e], -1.0 / 6.0);
    ad.k6[e] = 0.98 * pow(4.0 * ad.ct[e], -1.0 / 6.0);
  }
  return ad;
}

inline SchemeParams make_scheme_params(double eps, double ct) {
  SchemeParams s;
  s.eps = eps; s.teno_ct = ct; s.eps16 = 16.0 * eps;
  s.kfast5 = 0.98 * pow(3.0 * ct, -1.0 / 6.0);
  s.kfast6 = 0.98 * pow(4.0 * ct, -1.0 / 6.0);
  s.kpass5 = s.kfast5 - 1.0;
  s.slow_count = nullptr;
  return s;
}

OSB_HD double sq(double x) { return x * x; }

// max / min of two ordinary numbers as one compare + select.  fmax()/fmin() carry IEEE NaN semantics, which sm_100a has no
// FP64 instruction for: each call compiles to DSETP + ~8 integer/select instructions (cuobjdump, r02), and the sweeps are
// issue-bound.  The operands here are wave speeds / smoothness sums of a finite state.
OSB_HD double dmax2(double a, double b) { return a > b ? a : b; }
OSB_HD double dmin2(double a, double b) { return a < b ? a : b; }

// 1/x and 1/sqrt(x) for ordinary positive x: hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~23 bits) + two Newton steps, no
// special-case branches (the library routines guard denormals / infinities with a call to a slow path).  Relative error of the
// result <~ 2 ulp.
OSB_HD double rcp_nr(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}
OSB_HD double rsqrt_nr(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-(x * y), y, 1.0);
  y = fma(0.5 * y, e, y);
  e = fma(-(x * y), y, 1.0);
  return fma(0.5 * y, e, y);
#else
  return 1.0 / sqrt(x);
#endif
}
OSB_HD double pow6(double x) { double x2 = x * x; return x2 * x2 * x2; }

// 2^-e for m = 1.x * 2^e (m > 0, finite, normal): an exact power-of-two rescale that keeps the
// sixth powers of the TENO products inside the double range whatever eps / flux magnitudes are.
OSB_HD double inv_pow2(double m) {
#if defined(__CUDA_ARCH__)
  const long long b = __double_as_longlong(m);
  return __longlong_as_double((0x7FELL - ((b >> 52) & 0x7FFLL)) << 52);
#else
  long long b; memcpy(&b, &m, 8);
  b = (0x7FELL - ((b >> 52) & 0x7FFLL)) << 52;
  double r; memcpy(&r, &b, 8); return r;
#endif
}

// ------------------------------------------------------------------------------------------------
// One-sided reconstructions on f(-2..3) = f[0..5] of the right-biased split flux; the left-biased
// side calls the same function on the mirrored window (p -> 1-p).  The value returned is the
// reconstruction of 2*f (the 1/2 of the LLF splitting is applied by the caller through `half`).
// ------------------------------------------------------------------------------------------------
// TENO5 on g = 2f (doubled split flux); the contribution of one side to Recon is the reconstruction of f = g/2,
// sum_r omega_r q_r(f).  Split in a branch-free "front" (smoothness indicators, candidate polynomials, the cheap
// sufficient test for "every stencil passes the cut-off") and a "resolve" step, so that the two sides of a wave are
// scheduled together and share one branch.
//   beta''_r = 16 beta_r(f) = A_r^2 + 13/3 B_r^2   (teno.py:159-177 scaled; eps scaled alike, so tau/(eps+beta) is unchanged)
//   fast test: alpha_r >= 1 and sum(alpha) <= 3 alpha_max with alpha_max = (1 + tau/D_min)^6, hence
//   alpha_r/sum >= 1/(3 alpha_max) >= C_T whenever 1 + tau/D_min <= (3 C_T)^(-1/6)  (2 % margin in kfast5).
struct Teno5Side {
  double b0, b1, b2, tau;          // 16 beta_r(f) and tau_5 (before eps is added)
  double gm2, gm1, g0, g1, g2;
  bool all_pass;
};
OSB_HD Teno5Side teno5_front(double gm2, double gm1, double g0, double g1, double g2, const SchemeParams &sp) {
  Teno5Side t;
  const double e0 = gm1 - gm2, e1 = g0 - gm1, e2 = g1 - g0, e3 = g2 - g1;        // first differences
  const double A0 = e1 + e2, A1 = e3 - 3.0 * e2, A2 = 3.0 * e1 - e0;             // 2x first-derivative terms
  const double B0 = e2 - e1, B1 = e3 - e2, B2 = e1 - e0;                          // second differences
  const double c = 13.0 / 3.0;
  t.b0 = (c * B0) * B0 + A0 * A0; t.b1 = (c * B1) * B1 + A1 * A1; t.b2 = (c * B2) * B2 + A2 * A2;
  t.tau = fabs(t.b0 - t.b2);                                                      // teno.py:209
  t.gm2 = gm2; t.gm1 = gm1; t.g0 = g0; t.g1 = g1; t.g2 = g2;
  // sufficient test for "every stencil passes": 1 + tau/D_min <= kfast5  <=>  tau <= (kfast5 - 1)(eps + beta_r) for r = 0, 1, 2
  const double ke = sp.kpass5 * sp.eps16;
  t.all_pass = (t.tau <= fma(sp.kpass5, t.b0, ke)) & (t.tau <= fma(sp.kpass5, t.b1, ke)) & (t.tau <= fma(sp.kpass5, t.b2, ke));
  return t;
}
namespace teno5c {
constexpr double d0 = 11.0 / 20.0, d1 = 2.0 / 5.0, d2 = 1.0 / 20.0;               // teno.py:133
// 1/sum(d_r delta_r) takes one of 7 values (all-zero cannot happen: the largest alpha_r/sum is >= 1/3 > C_T);
// evaluated in the order of the run-time sum, hence bit-identical to the division.
constexpr double i111 = 1.0 / ((d0 + d1) + d2), i110 = 1.0 / ((d0 + d1) + 0.0), i101 = 1.0 / ((d0 + 0.0) + d2), i100 = 1.0 / ((d0 + 0.0) + 0.0);
constexpr double i011 = 1.0 / ((0.0 + d1) + d2), i010 = 1.0 / ((0.0 + d1) + 0.0), i001 = 1.0 / ((0.0 + 0.0) + d2);
}
// every stencil kept: the optimal weights give the 5-point linear upwind reconstruction of f = g/2
//   sum_r d_r q_r(g)/2 = (2 g(-2) - 18 g(-1) + 82 g(0) + 62 g(1) - 8 g(2)) / 240      (teno.py:113-133)
OSB_HD double teno5_linear(const Teno5Side &t) {
  return (1.0 / 120.0) * t.gm2 + ((-3.0 / 40.0) * t.gm1 + ((41.0 / 120.0) * t.g0 + ((31.0 / 120.0) * t.g1 + (-1.0 / 30.0) * t.g2)));
}
OSB_HD double teno5_resolve(const Teno5Side &t, const SchemeParams &sp) {
  using namespace teno5c;
  const double D0 = sp.eps16 + t.b0, D1 = sp.eps16 + t.b1, D2 = sp.eps16 + t.b2;
  // 12 x candidate reconstructions (teno.py:113-114 times 6, times 2 for g = 2f)
  const double Q0 = 5.0 * t.g0 + (2.0 * t.g1 - t.gm1);
  const double Q1 = 5.0 * t.g1 + (2.0 * t.g0 - t.g2);
  const double Q2 = 11.0 * t.g0 + (2.0 * t.gm2 - 7.0 * t.gm1);
  // P_r = (D_r + tau) * prod_{s != r} D_s ; alpha_r = (P_r / (D0 D1 D2))^6   (teno.py:207-212, C=1, q=6)
  const double P0 = (D0 + t.tau) * (D1 * D2), P1 = (D1 + t.tau) * (D0 * D2), P2 = (D2 + t.tau) * (D0 * D1);
  const double sc = inv_pow2(dmax2(P0, dmax2(P1, P2)));
  const double a0 = pow6(P0 * sc), a1 = pow6(P1 * sc), a2 = pow6(P2 * sc);
  const double thr = sp.teno_ct * (a0 + a1 + a2);                                 // teno.py:445-465
  const bool k0 = !(thr > a0), k1 = !(thr > a1), k2 = !(thr > a2);
  const double w0 = k0 ? (d0 / 12.0) : 0.0, w1 = k1 ? (d1 / 12.0) : 0.0, w2 = k2 ? (d2 / 12.0) : 0.0;
  const double inv = k0 ? (k1 ? (k2 ? i111 : i110) : (k2 ? i101 : i100)) : (k1 ? (k2 ? i011 : i010) : i001);
  return inv * (w0 * Q0 + w1 * Q1 + w2 * Q2);
}
// both sides of one characteristic wave: gp = CF + lam CS (right-biased), gm = CF - lam CS (left-biased, mirrored)
OSB_HD double teno5_wave(const double *gp, const double *gm, const SchemeParams &sp) {
  const Teno5Side tp = teno5_front(gp[0], gp[1], gp[2], gp[3], gp[4], sp);
  const Teno5Side tm = teno5_front(gm[5], gm[4], gm[3], gm[2], gm[1], sp);
  if (tp.all_pass && tm.all_pass) return teno5_linear(tp) + teno5_linear(tm);
  return teno5_resolve(tp, sp) + teno5_resolve(tm, sp);
}

// TENO6 (teno.py:85-136, 165-177, 216-234).  `square_last` reproduces the reference's right-biased
// beta_3 whose last term is not squared (teno.py:166-167).
OSB_HD double teno6_side(double fm2, double fm1, double f0, double f1, double f2, double f3,
                         const SchemeParams &sp, bool square_last) {
  const double b0 = 0.25 * sq(fm1 - f1) + (13.0 / 12.0) * sq(fm1 - 2.0 * f0 + f1);
  const double b1 = 0.25 * sq(3.0 * f0 - 4.0 * f1 + f2) + (13.0 / 12.0) * sq(f0 - 2.0 * f1 + f2);
  const double b2 = 0.25 * sq(fm2 - 4.0 * fm1 + 3.0 * f0) + (13.0 / 12.0) * sq(fm2 - 2.0 * fm1 + f0);
  const double l3 = -f0 + 3.0 * f1 - 3.0 * f2 + f3;
  const double b3 = (1.0 / 36.0) * sq(-11.0 * f0 + 18.0 * f1 - 9.0 * f2 + 2.0 * f3) +
                    (13.0 / 12.0) * sq(2.0 * f0 - 5.0 * f1 + 4.0 * f2 - f3) +
                    (781.0 / 720.0) * (square_last ? l3 * l3 : l3);
  const double tau = fabs(b3 - (1.0 / 6.0) * (b0 + b2 - 4.0 * b1));
  const double D0 = sp.eps + b0, D1 = sp.eps + b1, D2 = sp.eps + b2, D3 = sp.eps + b3;
  const double D01 = D0 * D1, D23 = D2 * D3;
  const double P0 = (D0 + tau) * (D1 * D23), P1 = (D1 + tau) * (D0 * D23);
  const double P2 = (D2 + tau) * (D01 * D3), P3 = (D3 + tau) * (D01 * D2);
  const double sc = inv_pow2(dmax2(dmax2(P0, P1), dmax2(P2, P3)));
  const double A0 = pow6(P0 * sc), A1 = pow6(P1 * sc), A2 = pow6(P2 * sc), A3 = pow6(P3 * sc);
  const double thr = sp.teno_ct * (A0 + A1 + A2 + A3);
  const double w0 = !(thr > A0) ? (231.0 / 500.0) : 0.0, w1 = !(thr > A1) ? (3.0 / 10.0) : 0.0;
  const double w2 = !(thr > A2) ? (27.0 / 500.0) : 0.0, w3 = !(thr > A3) ? (23.0 / 125.0) : 0.0;
  const double q0 = (1.0 / 6.0) * (-fm1 + 5.0 * f0 + 2.0 * f1);
  const double q1 = (1.0 / 6.0) * (2.0 * f0 + 5.0 * f1 - f2);
  const double q2 = (1.0 / 6.0) * (2.0 * fm2 - 7.0 * fm1 + 11.0 * f0);
  const double q3 = (1.0 / 12.0) * (3.0 * f0 + 13.0 * f1 - 5.0 * f2 + f3);
  const double inv = 1.0 / (w0 + w1 + w2 + w3);
  return inv * (w0 * q0 + w1 * q1 + w2 * q2 + w3 * q3);
}

// WENO5 JS / Z (weno.py:72-120 coefficients, 138-205 beta, 341-369 JS eps=1e-6 p=2, 283-338 Z eps=1e-14).
// NOTE the JS smoothness indicators are evaluated on f (= half of the argument g passed here), exactly as
// the reference evaluates them on 1/2(CF +- lambda CS): beta(f) = beta(g)/4.
template <bool Z>
OSB_HD double weno5_side(double gm2, double gm1, double g0, double g1, double g2) {
  const double b0 = 0.25 * ((13.0 / 12.0) * sq(g0 - 2.0 * g1 + g2) + 0.25 * sq(3.0 * g0 - 4.0 * g1 + g2));
  const double b1 = 0.25 * ((13.0 / 12.0) * sq(gm1 - 2.0 * g0 + g1) + 0.25 * sq(gm1 - g1));
  const double b2 = 0.25 * ((13.0 / 12.0) * sq(gm2 - 2.0 * gm1 + g0) + 0.25 * sq(gm2 - 4.0 * gm1 + 3.0 * g0));
  // alpha_r = d_r / (beta_r + eps)^2 (JS, weno.py:354) or d_r (1 + tau^2 / (beta_r + eps)^2) (Z, weno.py:323), each a division in
  // the reference.  Only the RATIOS alpha_r / sum(alpha) enter, so every alpha is taken over the common denominator
  // D0 D1 D2, D_r = (beta_r + eps)^2: n_r = d_r (D_r [+ tau^2 for Z] ... ) prod_{s != r} D_s -- one division per side instead
  // of four.  The D_r are rescaled by an exact power of two first, so the products stay inside the double range.
#ifdef OSB_WENO_FOUR_DIVISIONS        // the reference's form, kept for A/B measurements (scripts/weno_speed.py)
  {
    double a0, a1, a2;
    if (!Z) {
      const double e = 1.0e-6;
      a0 = (3.0 / 10.0) / sq(b0 + e); a1 = (3.0 / 5.0) / sq(b1 + e); a2 = (1.0 / 10.0) / sq(b2 + e);
    } else {
      const double e = 1.0e-14;
      const double t2 = sq(b0 - b2);
      a0 = (3.0 / 10.0) + (3.0 / 10.0) * t2 / sq(b0 + e);
      a1 = (3.0 / 5.0) + (3.0 / 5.0) * t2 / sq(b1 + e);
      a2 = (1.0 / 10.0) + (1.0 / 10.0) * t2 / sq(b2 + e);
    }
    const double q0 = (1.0 / 6.0) * (2.0 * g0 + 5.0 * g1 - g2);
    const double q1 = (1.0 / 6.0) * (-gm1 + 5.0 * g0 + 2.0 * g1);
    const double q2 = (1.0 / 6.0) * (2.0 * gm2 - 7.0 * gm1 + 11.0 * g0);
    return (a0 * q0 + a1 * q1 + a2 * q2) / (a0 + a1 + a2);
  }
#endif
  const double e = Z ? 1.0e-14 : 1.0e-6;
  double D0 = sq(b0 + e), D1 = sq(b1 + e), D2 = sq(b2 + e);
  const double sc = inv_pow2(dmax2(D0, dmax2(D1, D2)));
  D0 *= sc; D1 *= sc; D2 *= sc;
  double n0, n1, n2;
  if (!Z) {
    n0 = (3.0 / 10.0) * (D1 * D2); n1 = (3.0 / 5.0) * (D0 * D2); n2 = (1.0 / 10.0) * (D0 * D1);
  } else {
    const double t2 = sq(b0 - b2) * sc;
    n0 = (3.0 / 10.0) * ((D0 + t2) * (D1 * D2));
    n1 = (3.0 / 5.0) * ((D1 + t2) * (D0 * D2));
    n2 = (1.0 / 10.0) * ((D2 + t2) * (D0 * D1));
  }
  const double q0 = (1.0 / 6.0) * (2.0 * g0 + 5.0 * g1 - g2);
  const double q1 = (1.0 / 6.0) * (-gm1 + 5.0 * g0 + 2.0 * g1);
  const double q2 = (1.0 / 6.0) * (2.0 * gm2 - 7.0 * gm1 + 11.0 * g0);
  return (n0 * q0 + n1 * q1 + n2 * q2) / (n0 + n1 + n2);
}

// Both sides for one characteristic field.  cf[p], cs[p], p=0..5 <-> points -2..3; lam = max |lambda|.
// Returns Recon = recon+(1/2(CF + lam CS)) + recon-(1/2(CF - lam CS))   (shock_capturing.py:479-495).
template <int RECON>
OSB_HD double reconstruct(const double *cf, const double *cs, double lam, const SchemeParams &sp) {
  double gp[6], gm[6];
#pragma unroll
  for (int p = 0; p < 6; p++) { gp[p] = cf[p] + lam * cs[p]; gm[p] = cf[p] - lam * cs[p]; }
  double r;
  if (RECON == RECON_TENO5) {
    return teno5_wave(gp, gm, sp);
  } else if (RECON == RECON_TENO6) {
    // beta_3 of the right-biased side is not homogeneous (linear last term): evaluate on f itself.
    double fp[6], fm[6];
#pragma unroll
    for (int p = 0; p < 6; p++) { fp[p] = 0.5 * gp[p]; fm[p] = 0.5 * gm[p]; }
    return teno6_side(fp[0], fp[1], fp[2], fp[3], fp[4], fp[5], sp, false) +
           teno6_side(fm[5], fm[4], fm[3], fm[2], fm[1], fm[0], sp, true);
  } else if (RECON == RECON_WENO5_Z) {
    r = weno5_side<true>(gp[0], gp[1], gp[2], gp[3], gp[4]) + weno5_side<true>(gm[5], gm[4], gm[3], gm[2], gm[1]);
  } else {
    r = weno5_side<false>(gp[0], gp[1], gp[2], gp[3], gp[4]) + weno5_side<false>(gm[5], gm[4], gm[3], gm[2], gm[1]);
  }
  return 0.5 * r;
}

// Same, from the doubled split fluxes gp = CF + lam CS, gm = CF - lam CS already formed by the caller.
template <int RECON>
OSB_HD double reconstruct_g(const double *gp, const double *gm, const SchemeParams &sp) {
  if (RECON == RECON_TENO5) {
    return teno5_wave(gp, gm, sp);
  } else if (RECON == RECON_TENO6) {
    double fp[6], fm[6];
#pragma unroll
    for (int p = 0; p < 6; p++) { fp[p] = 0.5 * gp[p]; fm[p] = 0.5 * gm[p]; }
    return teno6_side(fp[0], fp[1], fp[2], fp[3], fp[4], fp[5], sp, false) +
           teno6_side(fm[5], fm[4], fm[3], fm[2], fm[1], fm[0], sp, true);
  } else if (RECON == RECON_WENO5_Z) {
    return 0.5 * (weno5_side<true>(gp[0], gp[1], gp[2], gp[3], gp[4]) + weno5_side<true>(gm[5], gm[4], gm[3], gm[2], gm[1]));
  } else {
    return 0.5 * (weno5_side<false>(gp[0], gp[1], gp[2], gp[3], gp[4]) + weno5_side<false>(gm[5], gm[4], gm[3], gm[2], gm[1]));
  }
}

// ------------------------------------------------------------------------------------------------
// Interface state and characteristic flux for one interface.
// Point data for the 6 stencil points p=0..5 (offsets -2..3 along the sweep direction DIR):
//   rho, m[ND] (momentum), E (rhoE), pr (pressure), a (speed of sound); velocities are m/rho.
// ------------------------------------------------------------------------------------------------
template <int ND>
struct Point {
  double rho, m[ND], E, pr, a, u[ND];
};

template <int ND, int DIR, int RECON, int AVG>
OSB_HD void interface_flux(const Point<ND> *pt, double gama, const SchemeParams &sp, double *flux) {
  constexpr int NV = ND + 2;
  const double gm1 = gama - 1.0;
  const Point<ND> &L = pt[2], &R = pt[3];
  // ---- interface state (averaging.py)
  double rho, u[ND], a;
  if (AVG == AVG_ROE) {
    const double sl = sqrt(L.rho), sr = sqrt(R.rho);
    rho = sqrt(L.rho * R.rho);
    const double w = 1.0 / (sr + sl);
    double ke = 0.0;
#pragma unroll
    for (int d = 0; d < ND; d++) { u[d] = w * (sr * R.u[d] + sl * L.u[d]); ke += u[d] * u[d]; }
    const double H = w * ((L.pr + L.E) / sl + (R.pr + R.E) / sr);
    a = sqrt(gm1 * (H - 0.5 * ke));
  } else {
    rho = 0.5 * (L.rho + R.rho);
#pragma unroll
    for (int d = 0; d < ND; d++) u[d] = 0.5 * (L.u[d] + R.u[d]);
    a = 0.5 * (L.a + R.a);
  }
  double ke2 = 0.0;
#pragma unroll
  for (int d = 0; d < ND; d++) ke2 += u[d] * u[d];

  double cf[NV][6], cs[NV][6], lam[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) lam[j] = 0.0;

  if (ND == 1) {
    // 1-D eigensystem in the reference's H-form (euler_eigensystem.py:57-75), ev = (u-a, u, u+a)
    const double ia = 1.0 / a, g = gm1 * ia * ia, ua = u[0] * ia;
    const double l00 = 0.25 * ua * (gm1 * ua + 2.0), l01 = -0.5 * ia * (gm1 * ua + 1.0), l02 = 0.5 * g;
    const double l10 = 1.0 - 0.5 * gm1 * ua * ua, l11 = g * u[0], l12 = -g;
    const double l20 = 0.25 * ua * (gm1 * ua - 2.0), l21 = -0.5 * ia * (gm1 * ua - 1.0), l22 = 0.5 * g;
#pragma unroll
    for (int p = 0; p < 6; p++) {
      const Point<ND> &P = pt[p];
      const double ud = P.u[0];
      const double F0 = P.m[0], F1 = P.m[0] * ud + P.pr, F2 = (P.pr + P.E) * ud;
      cf[0][p] = l00 * F0 + l01 * F1 + l02 * F2; cs[0][p] = l00 * P.rho + l01 * P.m[0] + l02 * P.E;
      cf[1][p] = l10 * F0 + l11 * F1 + l12 * F2; cs[1][p] = l10 * P.rho + l11 * P.m[0] + l12 * P.E;
      cf[2][p] = l20 * F0 + l21 * F1 + l22 * F2; cs[2][p] = l20 * P.rho + l21 * P.m[0] + l22 * P.E;
      lam[0] = fmax(lam[0], fabs(ud - P.a)); lam[1] = fmax(lam[1], fabs(ud)); lam[2] = fmax(lam[2], fabs(ud + P.a));
    }
  } else {
    // 2-D / 3-D (euler_eigensystem.py:77-135) with k~ = e_DIR:
    //   rows 0..ND-1 : row DIR = v0 - S/a^2 ; row r != DIR = -(e_DIR x w)_r / rho  (2-D: row 1 = (k1 w0 - k0 w1)/rho)
    //   row ND = beta (S + a w_DIR) ; row ND+1 = beta (S - a w_DIR) ;  beta = 1/(rho a sqrt2)
    const double phi = 0.5 * gm1 * ke2, ia2 = 1.0 / (a * a), irho = 1.0 / rho;
    const double bt = 0.70710678118654752440 * irho / a;
#pragma unroll
    for (int p = 0; p < 6; p++) {
      const Point<ND> &P = pt[p];
      const double ud = P.u[DIR];
      double v[2][NV];           // v[0] = q, v[1] = F
      v[0][0] = P.rho; v[1][0] = P.m[DIR];
#pragma unroll
      for (int d = 0; d < ND; d++) { v[0][1 + d] = P.m[d]; v[1][1 + d] = P.m[d] * ud + (d == DIR ? P.pr : 0.0); }
      v[0][ND + 1] = P.E; v[1][ND + 1] = (P.pr + P.E) * ud;
#pragma unroll
      for (int t = 0; t < 2; t++) {
        const double *x = v[t];
        double um = 0.0, w[ND];
#pragma unroll
        for (int d = 0; d < ND; d++) { um += u[d] * x[1 + d]; w[d] = x[1 + d] - u[d] * x[0]; }
        const double S = phi * x[0] - gm1 * um + gm1 * x[ND + 1];
        double c[NV];
        if (ND == 2) {
          c[0] = x[0] - S * ia2;
          c[1] = (DIR == 0 ? -w[1] : w[0]) * irho;      // (k1 w0 - k0 w1)/rho
        } else {
          // -(e_DIR x w)/rho : DIR=0 -> (.,  w2, -w1) ; DIR=1 -> (-w2, ., w0) ; DIR=2 -> (w1, -w0, .)
          const int r1 = (DIR + 1) % 3, r2 = (DIR + 2) % 3;
          c[DIR] = x[0] - S * ia2;
          c[r1] = w[r2 % ND] * irho;
          c[r2] = -w[r1 % ND] * irho;
        }
        const double aw = a * w[DIR];
        c[ND] = bt * (S + aw);
        c[ND + 1] = bt * (S - aw);
#pragma unroll
        for (int j = 0; j < NV; j++) { if (t == 0) cs[j][p] = c[j]; else cf[j][p] = c[j]; }
      }
      // max |lambda| over the stencil: ev = (U,..,U, U+a, U-a); repeated ones reuse the first
      lam[0] = fmax(lam[0], fabs(ud));
      lam[ND] = fmax(lam[ND], fabs(ud + P.a));
      lam[ND + 1] = fmax(lam[ND + 1], fabs(ud - P.a));
    }
#pragma unroll
    for (int j = 1; j < ND; j++) lam[j] = lam[0];
  }

  double rec[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) rec[j] = reconstruct<RECON>(cf[j], cs[j], lam[j], sp);

  // ---- flux = REV . rec
  if (ND == 1) {
    const double H = a * a / gm1 + 0.5 * u[0] * u[0];
    flux[0] = rec[0] + rec[1] + rec[2];
    flux[1] = (u[0] - a) * rec[0] + u[0] * rec[1] + (u[0] + a) * rec[2];
    flux[2] = (H - u[0] * a) * rec[0] + 0.5 * u[0] * u[0] * rec[1] + (H + u[0] * a) * rec[2];
  } else {
    const double al = 0.70710678118654752440 * rho / a;
    const double phig = 0.5 * ke2;                       // phi^2/(gama-1)
    const double Hp = phig + a * a / gm1;                // (phi^2 + a^2)/(gama-1)
    const double sp_ = al * (rec[ND] + rec[ND + 1]), sm = al * a * (rec[ND] - rec[ND + 1]);
    if (ND == 2) {
      const double k0 = DIR == 0 ? 1.0 : 0.0, k1 = DIR == 1 ? 1.0 : 0.0;
      flux[0] = rec[0] + sp_;
      flux[1] = u[0] * rec[0] + k1 * rho * rec[1] + u[0] * sp_ + k0 * sm;
      flux[2] = u[1] * rec[0] - k0 * rho * rec[1] + u[1] * sp_ + k1 * sm;
      flux[3] = phig * rec[0] + rho * (k1 * u[0] - k0 * u[1]) * rec[1] + Hp * sp_ + u[DIR] * sm;
    } else {
      // flux_rho = k.R3 + al(r3+r4); flux_m = (k.R3) u + rho (k x R3) + al[(r3+r4) u + a k (r3-r4)];
      // flux_E = (k.R3) phi^2/(g-1) + rho (u x k).R3 + al[(r3+r4) H' + theta a (r3-r4)]
      const int r1 = (DIR + 1) % 3, r2 = (DIR + 2) % 3;
      const double kR = rec[DIR];
      double fm_[3];
      fm_[DIR] = kR * u[DIR % ND] + u[DIR % ND] * sp_ + sm;
      fm_[r1] = kR * u[r1 % ND] - rho * rec[r2] + u[r1 % ND] * sp_;    // (e_DIR x R3)_{r1} = -R3_{r2}
      fm_[r2] = kR * u[r2 % ND] + rho * rec[r1] + u[r2 % ND] * sp_;    // (e_DIR x R3)_{r2} = +R3_{r1}
      flux[0] = kR + sp_;
      flux[1] = fm_[0]; flux[2] = fm_[1]; flux[3] = fm_[2];
      // (u x e_DIR).R3 = u_{r2} R3_{r1} - u_{r1} R3_{r2}
      flux[4] = kR * phig + rho * (u[r2 % ND] * rec[r1] - u[r1 % ND] * rec[r2]) + Hp * sp_ + u[DIR % ND] * sm;
    }
  }
}

// 4th-order central first / second derivative weights (scheme.py:81-85)
OSB_HD double d1c(double fm2, double fm1, double fp1, double fp2, double inv) {
  return (1.0 / 12.0) * inv * ((fm2 - fp2) + 8.0 * (fp1 - fm1));
}
OSB_HD double d2c(double fm2, double fm1, double f0, double fp1, double fp2, double inv2) {
  return (1.0 / 12.0) * inv2 * (16.0 * (fm1 + fp1) - (fm2 + fp2) - 30.0 * f0);
}

}  // namespace osb
