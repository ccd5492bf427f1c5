// osb_flux.cuh -- characteristic LLF flux of one interface from a *staged* 6-point window (v2).
//
// Same mathematics as shock_capturing.py:357-536 / euler_eigensystem.py:57-135 (see osb_math.cuh for the
// reconstructions), organised around what the FP64 pipe and the register file like:
//   * per staged point the kernel keeps  rho, 1/rho, m_k, E, p, a  in shared memory (constituent relations are
//     evaluated once per point while staging, not once per stencil use);
//   * one unified eigen-structure for 1-D/2-D/3-D:  entropy wave  c_E = v0 - S/a^2,
//     shear waves  c_t = +-w_t/rho^  (t != DIR, sign as in the reference's -(e_DIR x w)/rho^),
//     acoustic waves  c_+- = lsc (S +- a w_DIR), with
//       S = phi v0 - (g-1) u^.v_m + (g-1) v_E ,  w = v_m - u^ v0 ,  phi = (g-1)|u^|^2/2,
//       (lsc, rsc) = (1/(rho^ a sqrt2), rho^/(a sqrt2)) in 2-D/3-D and (1/(2a^2), 1) in 1-D  -- the reference's scalings;
//   * the characteristic FLUX is derived from the characteristic SOLUTION:  F = u_d q + p (0, e_d, u_d)  gives
//       CF_j = u_d CS_j + t_j ,   t_E = p (g-1)(u^_d - u_d)/a^2 ,  t_t = 0 ,  t_+- = lsc (+-a p - (g-1)(u^_d - u_d) p),
//     so each stencil point is projected once instead of twice;
//   * waves are processed one at a time; only S, w_DIR, u_d of the 6 points stay live between waves.
#pragma once
#include "osb_math.cuh"

namespace osb {

// index of the staged values
template <int ND> struct SV {
  static constexpr int RHO = 0, IRHO = 1, M0 = 2, E = 2 + ND, P = 3 + ND, A = 4 + ND, N = 5 + ND;
};

// sb points at value 0 of stencil point p=0 (offset -2); value v of point p is sb[v*VS + p*PS]
template <int ND, int DIR, int RECON, int AVG>
OSB_HD void interface_flux_staged(const double *sb, const int PS, const int VS, const double gama,
                                  const SchemeParams &sp, double *flux) {
  typedef SV<ND> V;
  constexpr int NV = ND + 2;
  const double gm1 = gama - 1.0;
#define SVAL(v, p) sb[(v) * VS + (p) * PS]
  // ---- interface state between points 2 and 3 (averaging.py:31-59 simple, 62-114 Roe)
  double rho, irho, u[ND], a, ia;
  {
    const double rL = SVAL(V::RHO, 2), rR = SVAL(V::RHO, 3), iL = SVAL(V::IRHO, 2), iR = SVAL(V::IRHO, 3);
    if (AVG == AVG_ROE) {
      const double sl = sqrt(rL), sr = sqrt(rR);
      rho = sl * sr;                                    // sqrt(rho_L rho_R)
      const double w = 1.0 / (sr + sl);
#pragma unroll
      for (int d = 0; d < ND; d++) u[d] = w * (sr * (SVAL(V::M0 + d, 3) * iR) + sl * (SVAL(V::M0 + d, 2) * iL));
      // (p+E)/sqrt(rho) = (p+E) * (1/rho) * sqrt(rho)
      const double H = w * ((SVAL(V::P, 2) + SVAL(V::E, 2)) * iL * sl + (SVAL(V::P, 3) + SVAL(V::E, 3)) * iR * sr);
      double ke = 0.0;
#pragma unroll
      for (int d = 0; d < ND; d++) ke += u[d] * u[d];
      a = sqrt(gm1 * (H - 0.5 * ke));
    } else {
      rho = 0.5 * (rL + rR);
#pragma unroll
      for (int d = 0; d < ND; d++) u[d] = 0.5 * (SVAL(V::M0 + d, 2) * iL + SVAL(V::M0 + d, 3) * iR);
      a = 0.5 * (SVAL(V::A, 2) + SVAL(V::A, 3));
    }
    irho = 1.0 / rho;
    ia = 1.0 / a;
  }
  double ke = 0.0;
#pragma unroll
  for (int d = 0; d < ND; d++) ke += u[d] * u[d];
  const double phi = 0.5 * gm1 * ke, ia2 = ia * ia;
  const double lsc = (ND == 1) ? 0.5 * ia2 : 0.70710678118654752440 * irho * ia;
  const double rsc = (ND == 1) ? 1.0 : 0.70710678118654752440 * rho * ia;

  // ---- pass 1 over the stencil: u_d, max wave speeds, S and w_DIR of the solution vector
  double ud[6], S[6], wd[6];
  double lam0 = 0.0, lamp = 0.0, lamm = 0.0;
#pragma unroll
  for (int p = 0; p < 6; p++) {
    const double r = SVAL(V::RHO, p), ap = SVAL(V::A, p);
    ud[p] = SVAL(V::M0 + DIR, p) * SVAL(V::IRHO, p);
    lam0 = fmax(lam0, fabs(ud[p]));
    lamp = fmax(lamp, fabs(ud[p] + ap));
    lamm = fmax(lamm, fabs(ud[p] - ap));
    double um = 0.0;
#pragma unroll
    for (int d = 0; d < ND; d++) um += u[d] * SVAL(V::M0 + d, p);
    S[p] = phi * r - gm1 * um + gm1 * SVAL(V::E, p);
    wd[p] = SVAL(V::M0 + DIR, p) - u[DIR] * r;
  }

  double gp[6], gm[6];
  // ---- entropy wave
  double recE;
  {
#pragma unroll
    for (int p = 0; p < 6; p++) {
      const double cs = SVAL(V::RHO, p) - S[p] * ia2;
      const double t = SVAL(V::P, p) * (gm1 * ia2) * (u[DIR] - ud[p]);
      gp[p] = (ud[p] + lam0) * cs + t;
      gm[p] = (ud[p] - lam0) * cs + t;
    }
    recE = reconstruct_g<RECON>(gp, gm, sp);
  }
  // ---- shear waves (tangential directions)
  double recT[ND > 1 ? ND : 1];
#pragma unroll
  for (int t = 0; t < ND; t++) {
    if (t == DIR) { recT[t] = 0.0; continue; }
    // reference sign convention -(e_DIR x w)_r / rho^ (matters for TENO6, whose beta_3 is not even in f)
    const double sg = (t == (DIR + 2) % 3) ? irho : -irho;
#pragma unroll
    for (int p = 0; p < 6; p++) {
      const double cs = (SVAL(V::M0 + t, p) - u[t] * SVAL(V::RHO, p)) * sg;
      gp[p] = (ud[p] + lam0) * cs;
      gm[p] = (ud[p] - lam0) * cs;
    }
    recT[t] = reconstruct_g<RECON>(gp, gm, sp);
  }
  // ---- acoustic waves u+a, u-a
  double recP, recM;
  {
#pragma unroll
    for (int p = 0; p < 6; p++) {
      const double pr = SVAL(V::P, p);
      const double e = gm1 * (u[DIR] - ud[p]) * pr;
      const double apr = a * pr, aw = a * wd[p];
      const double cs = lsc * (S[p] + aw), t = lsc * (apr - e);
      gp[p] = (ud[p] + lamp) * cs + t;
      gm[p] = (ud[p] - lamp) * cs + t;
    }
    recP = reconstruct_g<RECON>(gp, gm, sp);
#pragma unroll
    for (int p = 0; p < 6; p++) {
      const double pr = SVAL(V::P, p);
      const double e = gm1 * (u[DIR] - ud[p]) * pr;
      const double apr = a * pr, aw = a * wd[p];
      const double cs = lsc * (S[p] - aw), t = -lsc * (apr + e);
      gp[p] = (ud[p] + lamm) * cs + t;
      gm[p] = (ud[p] - lamm) * cs + t;
    }
    recM = reconstruct_g<RECON>(gp, gm, sp);
  }
#undef SVAL
  // ---- flux = REV . rec
  const double sp_ = rsc * (recP + recM), sm = rsc * a * (recP - recM);
  const double Hp = 0.5 * ke + a * a / gm1;
  flux[0] = recE + sp_;
  double fe = 0.5 * ke * recE + Hp * sp_ + u[DIR] * sm;
#pragma unroll
  for (int d = 0; d < ND; d++) {
    double fm_ = u[d] * recE + u[d] * sp_;
    if (d == DIR) fm_ += sm;
    else {
      const double sr = (d == (DIR + 2) % 3) ? rho : -rho;
      fm_ += sr * recT[d]; fe += sr * u[d] * recT[d];
    }
    flux[1 + d] = fm_;
  }
  flux[ND + 1] = fe;
}

}  // namespace osb
