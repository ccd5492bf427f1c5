// osb_flux3.cuh -- characteristic LLF flux sweeps, third generation: register-lean, occupancy-first.
//
// Same mathematics as osb_flux.cuh (shock_capturing.py:357-536, euler_eigensystem.py:57-135, teno.py:57-465, weno.py:35-465,
// averaging.py:31-114).  What changed is where the intermediate values live.  The second-generation kernels carried one
// interface per thread from the staged window to the flux in registers (168 registers, 12 warps per SM, FP64 pipe 50 %
// busy, top stall "wait" on dependent DFMA chains, ncu r01).  Here the work of an interface is cut into short passes that
// hand their results over through a THREAD-PRIVATE column of shared memory (no barrier: the same thread writes and reads):
//     pass 0   interface state (Roe / simple average), the three max wave speeds over the 6-point stencil
//     pass 1   S and w_d per stencil point -> doubled split fluxes g+ / g- of the two acoustic waves and S itself -> column
//     pass 2   rolled loop over the acoustic waves: reconstruction of both sides from the column (TENO5: 10 values per wave)
//     pass 3   rolled loop over the entropy wave (from S) and the shear waves (need neither S nor w_d): split fluxes formed
//              in registers from the staged window, reconstructed at once
//     pass 4   flux = R . rec ; parked in the thread's own column for the flux difference
// so that at most ~45 doubles are live at any point, and the rolled wave loops shrink the code (instruction-cache misses were
// the "no_instruction" stalls of the unrolled 5-wave body).  Round-2 ncu: the first cut of this design (all five waves through
// the column) ran into the shared-memory pipe (wavefronts 77 % of peak); only the waves that share S and w_d go through it now.
#pragma once
#include "osb_math.cuh"
#include "osb_flux.cuh"

namespace osb {

// values per wave handed to the reconstruction: TENO5 / WENO5 use f(-2..2) of each side, TENO6 all six.
// Column of a thread: two acoustic waves x NG split fluxes, then S of the 6 stencil points.
template <int RECON> struct F3 { static constexpr int NG = (RECON == RECON_TENO6) ? 12 : 10, NCOL = 2 * NG + 6; };

// both sides of one characteristic wave: a[] right-biased values (points 0..), b[] left-biased ones already mirrored
// (points 5, 4, ..); returns recon+(f+) + recon-(f-) with f+- = g+-/2  (shock_capturing.py:479-495)
template <int RECON>
OSB_HD double f3_recon(const double *a, const double *b, const SchemeParams &sp) {
  constexpr int NS = F3<RECON>::NG / 2;
  if (RECON == RECON_TENO5) {
    const Teno5Side tp = teno5_front(a[0], a[1], a[2], a[3], a[4], sp);
    const Teno5Side tm = teno5_front(b[0], b[1], b[2], b[3], b[4], sp);
    if (tp.all_pass && tm.all_pass) return teno5_linear(tp) + teno5_linear(tm);
#if defined(__CUDA_ARCH__)
    if (sp.slow_count) atomicAdd(sp.slow_count, 1ull);     // bench instrumentation: share of waves on the full cut-off path
#endif
    return teno5_resolve(tp, sp) + teno5_resolve(tm, sp);
  } else if (RECON == RECON_TENO6) {
    // beta_3 of the right-biased side is not homogeneous (linear last term, teno.py:166-167): evaluate on f = g/2 itself
    return teno6_side(0.5 * a[0], 0.5 * a[1], 0.5 * a[2], 0.5 * a[3], 0.5 * a[4], 0.5 * a[NS - 1], sp, false) +
           teno6_side(0.5 * b[0], 0.5 * b[1], 0.5 * b[2], 0.5 * b[3], 0.5 * b[4], 0.5 * b[NS - 1], sp, true);
  } else if (RECON == RECON_WENO5_Z) {
    return 0.5 * (weno5_side<true>(a[0], a[1], a[2], a[3], a[4]) + weno5_side<true>(b[0], b[1], b[2], b[3], b[4]));
  } else {
    return 0.5 * (weno5_side<false>(a[0], a[1], a[2], a[3], a[4]) + weno5_side<false>(b[0], b[1], b[2], b[3], b[4]));
  }
}

// store g+ of stencil point p (slot p) and g- of point p mirrored (slot NS + 5 - p) if the reconstruction reads them
template <int RECON>
OSB_HD void f3_put(double *G, const int GS, const int p, const double gp, const double gm) {
  constexpr int NS = F3<RECON>::NG / 2;
  if (p < NS) G[p * GS] = gp;
  if (5 - p < NS) G[(NS + 5 - p) * GS] = gm;
}

// sb: staged value 0 of stencil point 0 (offset -2); value v of point p at sb[v*VS + p*PS].
// G: this thread's column (element k at G[k*GS]), NW*NG doubles.  flux: ND+2 values out.
// where the 6 stencil points of an interface sit in the staged array: equally spaced (tiles), or rows of a ring buffer of
// 2^k rows of 32 lanes (marching kernel: the window may straddle the wrap)
struct WinAffine { int PS; OSB_HD int off(int p) const { return p * PS; } };
struct WinRing { int row0, mask; OSB_HD int off(int p) const { return ((row0 + p) & mask) * 32; } };

#ifndef OSB_F3_UNROLL_ACOUSTIC
#define OSB_F3_UNROLL_ACOUSTIC 1     // 2: both acoustic waves interleaved (more ILP, more registers)
#endif
constexpr int F3_UNROLL_ACOUSTIC = OSB_F3_UNROLL_ACOUSTIC;
template <int ND, int DIR, int RECON, int AVG, typename WIN>
OSB_HD void interface_flux_split(const double *sb, const WIN win, const int VS, double *G, const int GS,
                                 const double gama, const SchemeParams &sp, double *flux) {
  typedef SV<ND> V;
  constexpr int NG = F3<RECON>::NG, NS = NG / 2;
  const double gm1 = gama - 1.0;
#define SVAL(v, p) sb[(v) * VS + win.off(p)]
  // ---- pass 0: interface state between points 2 and 3 (averaging.py:31-59 simple, 62-114 Roe)
  double rho, irho, u[ND], a, ia, hst;               // hst = a^2/(gama-1)
  {
    const double rL = SVAL(V::RHO, 2), rR = SVAL(V::RHO, 3), yL = SVAL(V::Y, 2), yR = SVAL(V::Y, 3);
    if (AVG == AVG_ROE) {
      const double sl = rL * yL, sr = rR * yR;          // sqrt(rho_L), sqrt(rho_R)
      rho = sl * sr;
      irho = yL * yR;
      const double w = rcp_nr(sr + sl);
      // sqrt(rho) u = m y  and  (p + E)/sqrt(rho) = (p + E) y
#pragma unroll
      for (int d = 0; d < ND; d++) u[d] = w * (SVAL(V::M0 + d, 3) * yR + SVAL(V::M0 + d, 2) * yL);
      const double H = w * ((SVAL(V::P, 2) + SVAL(V::E, 2)) * yL + (SVAL(V::P, 3) + SVAL(V::E, 3)) * yR);
      double ke = 0.0;
#pragma unroll
      for (int d = 0; d < ND; d++) ke += u[d] * u[d];
      hst = H - 0.5 * ke;
      const double a2 = gm1 * hst;
      ia = rsqrt_nr(a2);
      a = a2 * ia;
    } else {
      rho = 0.5 * (rL + rR);
#pragma unroll
      for (int d = 0; d < ND; d++) u[d] = 0.5 * (SVAL(V::M0 + d, 2) * (yL * yL) + SVAL(V::M0 + d, 3) * (yR * yR));
      a = 0.5 * (SVAL(V::A, 2) + SVAL(V::A, 3));
      irho = rcp_nr(rho);
      ia = rcp_nr(a);
      hst = a * a * rcp_nr(gm1);
    }
  }
  double ke = 0.0;
#pragma unroll
  for (int d = 0; d < ND; d++) ke += u[d] * u[d];
  const double phi = 0.5 * gm1 * ke, ia2 = ia * ia;
  const double lsc = (ND == 1) ? 0.5 * ia2 : 0.70710678118654752440 * irho * ia;
  const double rsc = (ND == 1) ? 1.0 : 0.70710678118654752440 * rho * ia;

  // local wave speeds: max over the stencil of |u_d|, |u_d + a|, |u_d - a|  (shock_capturing.py:512-536)
  double lam0 = 0.0, lamp = 0.0, lamm = 0.0;
#pragma unroll
  for (int p = 0; p < 6; p++) {
    const double ap = SVAL(V::A, p), udp = SVAL(V::UD, p);
    lam0 = dmax2(lam0, fabs(udp));
    lamp = dmax2(lamp, fabs(udp + ap));
    lamm = dmax2(lamm, fabs(udp - ap));
  }
  // ---- pass 1: acoustic split fluxes and S -> column
  // split fluxes g+- = CF +- lam CS (shock_capturing.py:479-495; doubled, the 1/2 is in the reconstruction) with
  //   entropy   CS = rho - S/a^2          CF = u_d CS + e/a^2          e = (g-1)(u^_d - u_d) p
  //   shear     CS = +-(m_t - u^_t rho)/rho^    CF = u_d CS
  //   acoustic  CS = lsc (S +- a w_d)     CF = u_d CS + lsc (+-a p - e)
  const double la = lsc * a;
  double *const colS = G + 2 * NG * GS;
#pragma unroll
  for (int p = 0; p < 6; p++) {
    const double r = SVAL(V::RHO, p), pr = SVAL(V::P, p), udp = SVAL(V::UD, p);
    double um = 0.0;
#pragma unroll
    for (int d = 0; d < ND; d++) um = fma(u[d], SVAL(V::M0 + d, p), um);
    const double S = fma(phi, r, gm1 * (SVAL(V::E, p) - um));
    const double wd = fma(-u[DIR], r, SVAL(V::M0 + DIR, p));
    const double e = (gm1 * (u[DIR] - udp)) * pr;
    colS[p * GS] = S;
    const double lS = lsc * S, lw = la * wd, lp = la * pr, le = lsc * e;
    {
      const double cs = lS + lw;
      const double cf = fma(udp, cs, lp - le);
      f3_put<RECON>(G, GS, p, fma(lamp, cs, cf), fma(-lamp, cs, cf));
    }
    {
      const double cs = lS - lw;
      const double cf = fma(udp, cs, -(lp + le));
      f3_put<RECON>(G + NG * GS, GS, p, fma(lamm, cs, cf), fma(-lamm, cs, cf));
    }
  }
  // ---- pass 2: the acoustic reconstructions, one wave at a time (rolled: one copy of the reconstruction code)
#pragma unroll F3_UNROLL_ACOUSTIC
  for (int w = 0; w < 2; w++) {
    double ga[NS], gb[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) { ga[k] = G[(w * NG + k) * GS]; gb[k] = G[(w * NG + NS + k) * GS]; }
    G[w * NG * GS] = f3_recon<RECON>(ga, gb, sp);
  }
  // ---- pass 3: entropy wave (k = 0) and shear waves (k = 1 .. ND-1: the tangential directions in ascending order), split
  // fluxes formed in registers
#pragma unroll 1
  for (int k = 0; k < ND; k++) {
    double ga[NS], gb[NS];
    if (k == 0) {
#pragma unroll
      for (int p = 0; p < 6; p++) {
        const double udp = SVAL(V::UD, p);
        const double e = (gm1 * (u[DIR] - udp)) * SVAL(V::P, p);
        const double cs = fma(-colS[p * GS], ia2, SVAL(V::RHO, p));
        const double cf = fma(udp, cs, e * ia2);
        if (p < NS) ga[p] = fma(lam0, cs, cf);
        if (5 - p < NS) gb[5 - p] = fma(-lam0, cs, cf);
      }
    } else {
      // tangential direction t of wave k; reference sign convention -(e_DIR x w)_t / rho^ (matters for TENO6, whose beta_3
      // is not even in f)
      const int t = (k - 1 < DIR) ? k - 1 : k;
      const double ut = (ND > 1 && t == 0) ? u[0] : ((ND > 2 && t == 2) ? u[ND > 2 ? 2 : 0] : u[ND > 1 ? 1 : 0]);
      const double sg = (t == (DIR + 2) % 3) ? irho : -irho;
#pragma unroll
      for (int p = 0; p < 6; p++) {
        const double cs = (SVAL(V::M0 + t, p) - ut * SVAL(V::RHO, p)) * sg;
        const double cf = SVAL(V::UD, p) * cs;
        if (p < NS) ga[p] = fma(lam0, cs, cf);
        if (5 - p < NS) gb[5 - p] = fma(-lam0, cs, cf);
      }
    }
    colS[k * GS] = f3_recon<RECON>(ga, gb, sp);           // S is not needed any more after k = 0 has read it
  }
  const double recP = G[0], recM = G[NG * GS], recE = colS[0];
  double recT[ND > 1 ? ND : 1];
  if (ND > 1) {
#pragma unroll
    for (int t = 0; t < ND; t++) recT[t] = (t == DIR) ? 0.0 : colS[(t < DIR ? t + 1 : t) * GS];
  }
#undef SVAL
  // ---- pass 4: flux = REV . rec
  const double sp_ = rsc * (recP + recM), sm = rsc * a * (recP - recM);
  const double Hp = 0.5 * ke + hst;
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

#if defined(__CUDACC__)
#include "osb_types.cuh"
namespace osb {

// -------------------------------------------------------------------------------------------------
// Sweep kernels.  Points are numbered along the sweep direction including 3 halo points on both sides (-3 .. np+2) and
// pencils are concatenated, so a block is a run of consecutive staged points (x) or staged rows (y/z); consecutive blocks
// overlap by 6.  One interface per thread; the thread's column G first holds the split fluxes, then its 5 flux values for
// the flux difference (read by the thread of the next point after a barrier).
// Shared memory per block:  x   : 128 points x (NVAL staged + NW*NG column) doubles      = 38 KB (TENO5, 3-D) -> 5 blocks / SM
//                           y/z : (TY+5) rows x 32 x NVAL + 32*TY x NW*NG doubles        = 107 KB (TY = 10)   -> 2 blocks / SM
// -------------------------------------------------------------------------------------------------
constexpr int F3_BT = 128;                 // x sweep: staged points (= threads) per block, F3_BT - 6 residual points
#ifndef OSB_F3_TY
#define OSB_F3_TY 10
#endif
template <int RECON> struct F3TY { static constexpr int v = RECON == RECON_TENO6 ? 7 : OSB_F3_TY; };
template <int RECON> constexpr int f3_ty() { return F3TY<RECON>::v; }     // y/z sweeps: interface rows (= thread rows) per block
template <int ND, int RECON> constexpr size_t f3_x_smem_bytes() { return sizeof(double) * F3_BT * (SV<ND>::N + F3<RECON>::NCOL); }
template <int ND, int RECON> constexpr size_t f3_yz_smem_bytes() {
  return sizeof(double) * 32 * ((f3_ty<RECON>() + 5) * SV<ND>::N + f3_ty<RECON>() * F3<RECON>::NCOL);
}

#ifndef OSB_F3_XBLOCKS
#define OSB_F3_XBLOCKS 5
#endif
#ifndef OSB_F3_YZBLOCKS
#define OSB_F3_YZBLOCKS 2
#endif

template <int ND, int RECON, int AVG, bool ACCUM>
__global__ void __launch_bounds__(F3_BT, OSB_F3_XBLOCKS) k_flux3_x(GridDev g, FieldPtrs f, PhysConst c, SchemeParams sp, AdaptiveCT ad, GeneralPtrs gp, SweepIdx si) {
  constexpr int NV = ND + 2, NVAL = SV<ND>::N;
  extern __shared__ double f3_smem[];
  double *sP = f3_smem;                          // [NVAL][F3_BT]
  double *sG = f3_smem + NVAL * F3_BT;           // [NW*NG][F3_BT]
  const int t = threadIdx.x;
  const unsigned fidx = blockIdx.x * (unsigned)(F3_BT - 6) + t;
  const bool inside = fidx < si.total;
  int ip = 0;
  long long x = 0;
  if (inside) {
    const unsigned row = si.len.div(fidx);
    ip = (int)(fidx - row * si.len.d) - 3;
    const unsigned k = ND > 2 ? si.n1.div(row) : 0u, j = row - k * si.n1.d;
    x = g.off + ip + (ND > 1 ? j * g.s[1] : 0) + (ND > 2 ? k * g.s[2] : 0);
    stage_point<ND, 0>(f, x, c.gama, sP + t, F3_BT);
  }
  __syncthreads();
  const bool iface = inside && t >= 2 && t <= F3_BT - 4 && ip >= -1 && ip <= g.np[0] - 1;
  const bool point = iface && t >= 3 && ip >= 0;
  if (ACCUM && point && (t & 3) == 0) {          // old Residual lines requested into L2 now, loaded after the reconstruction
#pragma unroll
    for (int m = 0; m < NV; m++) prefetch_l2(f.R[m] + x);
  }
  double fl[NV];
#pragma unroll
  for (int m = 0; m < NV; m++) fl[m] = 0.0;
  if (iface) {
    if (ad.on) {                                 // sensor value of the interface's left point (0 in the halos)
      const int e = adaptive_exponent(ad, gp.theta[x]);
      sp.teno_ct = ad.ct[e]; sp.kfast5 = ad.k5[e]; sp.kpass5 = ad.k5[e] - 1.0; sp.kfast6 = ad.k6[e];
      if (gp.teno_store) gp.teno_store[x] = sp.teno_ct;
    }
    interface_flux_split<ND, 0, RECON, AVG>(sP + t - 2, WinAffine{1}, F3_BT, sG + t, F3_BT, c.gama, sp, fl);
  }
  // flux difference along x, the fastest axis: the flux of the lower interface comes from the neighbouring lane by a warp
  // shuffle; only the last lane of a warp parks its flux in shared memory for the first lane of the next warp
  const int lane = t & 31;
  double lower[NV];
#pragma unroll
  for (int m = 0; m < NV; m++) lower[m] = __shfl_up_sync(0xffffffffu, fl[m], 1);
  if (lane == 31) {
#pragma unroll
    for (int m = 0; m < NV; m++) sG[m * F3_BT + t] = fl[m];
  }
  __syncthreads();
  if (point) {
    const double met = gp.D[0] ? -c.inv[0] * gp.D[0][x] : -c.inv[0];
    double old[NV];
    if (ACCUM) {                                 // all loads before the first store (the Residual arrays may alias for the compiler)
#pragma unroll
      for (int m = 0; m < NV; m++) old[m] = f.R[m][x];
    }
#pragma unroll
    for (int m = 0; m < NV; m++) {
      const double lo = lane == 0 ? sG[m * F3_BT + t - 1] : lower[m];
      const double r = met * (fl[m] - lo);
      f.R[m][x] = ACCUM ? old[m] + r : r;
    }
  }
}

template <int ND, int DIR, int RECON, int AVG, bool ACCUM>
__global__ void __launch_bounds__(32 * F3TY<RECON>::v, OSB_F3_YZBLOCKS) k_flux3_yz(GridDev g, FieldPtrs f, PhysConst c, SchemeParams sp, AdaptiveCT ad, GeneralPtrs gp, SweepIdx si) {
  constexpr int NV = ND + 2, NVAL = SV<ND>::N;
  constexpr int TY = F3TY<RECON>::v, RT = TY + 5, NTH = 32 * TY;
  constexpr int OTH = (DIR == 1) ? 2 : 1;
  extern __shared__ double f3_smem[];
  double *sP = f3_smem;                          // [NVAL][RT][32]
  double *sG = f3_smem + NVAL * RT * 32;         // [NW*NG][TY][32]
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const int i = blockIdx.y * 32 + tx;
  const bool xin = i < g.np[0];
  const unsigned r0 = blockIdx.x * (unsigned)(TY - 1);
  auto locate = [&](int r, int &jp, long long &x) -> bool {       // staged row r of this block -> grid index along DIR, linear index
    const unsigned fr = r0 + r;
    if (!xin || fr >= si.total) return false;
    const unsigned o = si.len.div(fr);
    jp = (int)(fr - o * si.len.d) - 3;
    x = g.off + i + (long long)jp * g.s[DIR] + (ND > 2 ? o * g.s[OTH] : 0);
    return true;
  };
  // stage RT rows (all loads issued before the first use)
  {
    constexpr int NIT = (RT + TY - 1) / TY;
    double raw[NIT][NV];
    bool ok[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const int r = ty + it * TY;
      int jp; long long x;
      ok[it] = r < RT && locate(r, jp, x);
      if (ok[it]) {
#pragma unroll
        for (int m = 0; m < NV; m++) raw[it][m] = __ldg(f.q[m] + x);
      }
    }
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const int r = ty + it * TY;
      if (ok[it]) stage_values<ND, DIR>(raw[it], c.gama, sP + r * 32 + tx, RT * 32);
    }
  }
  __syncthreads();
  // interface row r = 2 + ty (interface between staged rows r and r+1); its point row is r for r >= 3
  const int r = 2 + ty;
  int jp = 0; long long x = 0;
  const bool live = locate(r, jp, x);
  const bool iface = live && jp >= -1 && jp <= g.np[DIR] - 1;
  const bool point = iface && ty >= 1 && jp >= 0;
  if (ACCUM && point && (tx & 3) == 0) {
#pragma unroll
    for (int m = 0; m < NV; m++) prefetch_l2(f.R[m] + x);
  }
  if (iface) {
    double fl[NV];
    if (ad.on) {
      const int e = adaptive_exponent(ad, gp.theta[x]);
      sp.teno_ct = ad.ct[e]; sp.kfast5 = ad.k5[e]; sp.kpass5 = ad.k5[e] - 1.0; sp.kfast6 = ad.k6[e];
    }
    interface_flux_split<ND, DIR, RECON, AVG>(sP + (r - 2) * 32 + tx, WinAffine{32}, RT * 32, sG + tid, NTH, c.gama, sp, fl);
#pragma unroll
    for (int m = 0; m < NV; m++) sG[m * NTH + tid] = fl[m];
  }
  __syncthreads();
  if (point) {
    const double met = gp.D[DIR] ? -c.inv[DIR] * gp.D[DIR][x] : -c.inv[DIR];
    double old[NV];
    if (ACCUM) {
#pragma unroll
      for (int m = 0; m < NV; m++) old[m] = f.R[m][x];
    }
#pragma unroll
    for (int m = 0; m < NV; m++) {
      const double rr = met * (sG[m * NTH + tid] - sG[m * NTH + tid - 32]);
      f.R[m][x] = ACCUM ? old[m] + rr : rr;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// y / z sweeps of 3-D boxes as a MARCH along the sweep direction with a TMA + mbarrier pipeline.
// A block owns 32 x-lanes of one pencil (fixed index in the other direction) and a chunk of rows along DIR; it advances
// MTY interface rows per step.  Rows are staged once: a ring of 2 groups x MTY staged rows stays in shared memory, so the
// tile kernel's re-staging (15 staged rows per 9 point rows) and its recomputed interface row disappear, and the loads of
// the next row group are in flight while the current one is reconstructed:
//     one elected thread:  mbarrier.arrive.expect_tx + 5 x cp.async.bulk.tensor.3d (box 32 x MTY rows of rho .. rhoE) -> raw
//     all threads:         reconstruct the MTY interface rows of group n from the ring ; flux difference -> Residual
//                          mbarrier.try_wait (group n+2 has landed) ; constituent relations raw -> ring slot of group n
// Needs an even padded x-extent (TMA global strides are multiples of 16 bytes) and enough pencils to fill the GPU; the
// tile kernel k_flux3_yz covers everything else.
// -------------------------------------------------------------------------------------------------
#ifndef OSB_F3_MTY
#define OSB_F3_MTY 8
#endif
constexpr int F3_MTY = OSB_F3_MTY;                 // interface rows (= thread rows) per step; ring = 2 * MTY rows (power of two)
static_assert((F3_MTY & (F3_MTY - 1)) == 0, "ring addressing needs a power-of-two row group");
// The box starts one element left of the block's first lane: TMA wants the start of a box 16-byte aligned along x (probed on
// the B200: an odd start coordinate of 8-byte elements raises "illegal instruction"), and the interior starts at the odd
// padded index 5.  34 doubles per row = 272 bytes (a multiple of 16), rows of the landed box are 4 banks apart.
constexpr int F3_MBOX = 34;
template <int RECON> constexpr size_t f3_march_smem_bytes() {
  return sizeof(double) * (5 * F3_MTY * F3_MBOX + 32 * (2 * F3_MTY * SV<3>::N + F3_MTY * F3<RECON>::NCOL + 2 * 5)) + 128;
}

template <int DIR, int RECON, int AVG, bool ACCUM>
__global__ void __launch_bounds__(32 * F3_MTY, 2) k_flux3_march(GridDev g, FieldPtrs f, PhysConst c, SchemeParams sp,
                                                                 const __grid_constant__ TmaMaps5 maps, int chunk) {
  constexpr int ND = 3, NV = 5, NVAL = SV<3>::N, TY = F3_MTY, RING = 2 * TY, NTH = 32 * TY;
  constexpr int OTH = (DIR == 1) ? 2 : 1;
  extern __shared__ __align__(128) double f3m_smem[];
  double *raw = f3m_smem;                          // [NV][TY][34]   TMA destination (one row group, lanes -1 .. 32)
  double *ring = raw + NV * TY * F3_MBOX;          // [NVAL][RING][32]
  double *sG = ring + NVAL * RING * 32;            // [NCOL][TY][32] thread-private columns, then the fluxes
  double *prevF = sG + F3<RECON>::NCOL * NTH;      // [2][NV][32]    flux of the last interface row of the previous step
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(prevF + 2 * NV * 32);
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const int i0 = blockIdx.x * 32, i = i0 + tx, o = blockIdx.y;
  const bool xin = i < g.np[0];
  // rows of this chunk: points [p0, p1) along DIR, interfaces jp = p0-1 .. p1-1; local staged row ls <-> grid index p0 + ls - 3
  const int p0 = blockIdx.z * chunk, p1 = min(p0 + chunk, g.np[DIR]);
  const int nsteps = (p1 - p0 + 1 + TY - 1) / TY;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int group) {                    // thread 0: row group `group` (staged rows group*TY ..) -> raw
    mbar_expect_tx(bar, NV * TY * F3_MBOX * (unsigned)sizeof(double));
    const int row = p0 + group * TY - 3 + g.h;     // padded index of the group's first row
#pragma unroll
    for (int m = 0; m < NV; m++) {
      if (DIR == 1) tma_load_3d(raw + m * TY * F3_MBOX, maps.m[m], bar, i0 + g.h - 1, row, o + g.h);
      else tma_load_3d(raw + m * TY * F3_MBOX, maps.m[m], bar, i0 + g.h - 1, o + g.h, row);
    }
  };
  auto convert = [&](int group) {                  // constituent relations of this thread's point of the landed group
    double q[NV];
#pragma unroll
    for (int m = 0; m < NV; m++) q[m] = raw[(m * TY + ty) * F3_MBOX + 1 + tx];
    stage_values<ND, DIR>(q, c.gama, ring + ((group & 1) * TY + ty) * 32 + tx, RING * 32);
  };
  unsigned phase = 0;
  if (tid == 0) issue(0);
  mbar_wait(bar, phase); phase ^= 1;
  convert(0);
  __syncthreads();
  if (tid == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(1); }
  mbar_wait(bar, phase); phase ^= 1;
  convert(1);
  __syncthreads();
  if (tid == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(2); }
  const long long xbase = g.off + i + (long long)o * g.s[OTH];
#pragma unroll 1
  for (int n = 0; n < nsteps; n++) {
    const int ls = 2 + n * TY + ty;                // local staged row left of this thread's interface
    const int jp = p0 + ls - 3;
    const bool iface = xin && jp <= p1 - 1;        // jp >= p0 - 1 >= -1 always
    const bool point = iface && ls >= 3;           // lower interface ls-1 belongs to this chunk too
    const long long x = xbase + (long long)jp * g.s[DIR];
    if (ACCUM && point && (tx & 3) == 0) {
#pragma unroll
      for (int m = 0; m < NV; m++) prefetch_l2(f.R[m] + x);
    }
    if (iface) {
      double fl[NV];
      interface_flux_split<ND, DIR, RECON, AVG>(ring + tx, WinRing{ls - 2, RING - 1}, RING * 32, sG + tid, NTH, c.gama, sp, fl);
#pragma unroll
      for (int m = 0; m < NV; m++) sG[m * NTH + tid] = fl[m];
      if (ty == TY - 1) {
#pragma unroll
        for (int m = 0; m < NV; m++) prevF[((n & 1) * NV + m) * 32 + tx] = fl[m];
      }
    }
    __syncthreads();                               // fluxes of the step are in the columns; ring group n is no longer read
    if (point) {
      const double met = -c.inv[DIR];
      double old[NV];
      if (ACCUM) {
#pragma unroll
        for (int m = 0; m < NV; m++) old[m] = f.R[m][x];
      }
#pragma unroll
      for (int m = 0; m < NV; m++) {
        const double lower = ty > 0 ? sG[m * NTH + tid - 32] : prevF[(((n + 1) & 1) * NV + m) * 32 + tx];
        const double rr = met * (sG[m * NTH + tid] - lower);
        f.R[m][x] = ACCUM ? old[m] + rr : rr;
      }
    }
    if (n + 1 < nsteps) {                          // group n+2 replaces group n in the ring; group n+3 starts travelling
      mbar_wait(bar, phase); phase ^= 1;
      convert(n + 2);
      __syncthreads();
      if (tid == 0 && n + 2 < nsteps) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(n + 3); }
    }
  }
}

}  // namespace osb
#endif  // __CUDACC__
