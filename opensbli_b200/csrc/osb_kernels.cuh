// osb_kernels.cuh -- sm_100a kernels of the per-timestep solver loop (fp64).
//
// Kernel families (OSB_FAM_* in include/osbli_b200.h) and the reference loops they replace
// (SURVEY.md section 2b / appendix A):
//   k_prim                     CRu_i, CRp, CRa, CRT, CRmu            (constituent relations, one fused pass)
//   k_theta                    modified Ducros sensor (adaptive TENO)
//   k_flux3_x / k_flux3_yz     LLF{Weno,Teno}_reconstruction_d + Residual (staged flux sweep + flux difference, fused): osb_flux3.cuh
//   k_flux_curv2d / k_resid_curv2d   the same on fully curvilinear 2-D grids (metric-aware eigensystem)
//   k_central / k_central_general    Convective terms group / CD / residual (Blaisdell / Feiereisen splits, closures, metrics)
//   k_viscous / k_viscous_general    Derivative evaluation CD + Viscous terms (fused, mixed derivatives on the fly)
//   k_viscous3d_tiled          3-D viscous terms + RK update (+ constituent relations from q, + peer stores of a slab run)
//   k_viscous3d_tiled_general  3-D viscous terms with variable viscosity / metrics / closures
//   k_central3d_fused          whole Central(4) stage in one kernel
//   k_rk_*                     Save equations / Sub stage / Temporal solution advancement
//   k_copy_box, k_fill_box, k_bc_*   exchange{n}_block0 and the boundary-condition kernels
//   k_signal / k_wait          stream-ordered handshakes between the ranks of a slab decomposition
#pragma once
#include "osb_math.cuh"
#include "osb_flux.cuh"
#include "osb_types.cuh"

namespace osb {

// -------------------------------------------------------------------------------------------------
// constituent relations over [lo, hi) (grid + scheme halos)
// -------------------------------------------------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(256) k_prim(GridDev g, FieldPtrs f, PhysConst c, double *mu, int lo0, int lo1, int lo2, int n0, int n1, int n2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= n0 || j >= n1 || k >= n2) return;
  const long long x = g.off + (lo0 + i) + (ND > 1 ? (lo1 + j) * g.s[1] : 0) + (ND > 2 ? (lo2 + k) * g.s[2] : 0);
  // all loads first: the output arrays may alias the inputs as far as the compiler knows
  double q[ND + 2];
#pragma unroll
  for (int m = 0; m < ND + 2; m++) q[m] = __ldg(f.q[m] + x);
  const double rho = q[0];
  double ke = 0.0, u[ND];
#pragma unroll
  for (int d = 0; d < ND; d++) {
    u[d] = q[1 + d] / rho;
    ke += 0.5 * rho * u[d] * u[d];
  }
  const double p = (c.gama - 1.0) * (q[ND + 1] - ke);
#pragma unroll
  for (int d = 0; d < ND; d++) f.u[d][x] = u[d];
  f.p[x] = p;
  f.a[x] = sqrt(c.gama * p / rho);
  const double T = c.Minf * c.Minf * c.gama * p / rho;
  f.T[x] = T;
  if (mu) {   // viscosity laws of the apps: Sutherland (katzer_SBLI.py:25), power law (turbulent_channel.py:29)
    if (c.visc_law == 1) mu[x] = (c.SuthT / c.RefT + 1.0) * (T * sqrt(T)) / (c.SuthT / c.RefT + T);
    else if (c.visc_law == 2) mu[x] = pow(T, c.mu_exp);
    else mu[x] = 1.0;
  }
}

// -------------------------------------------------------------------------------------------------
// Central(4) skew-symmetric convective residual (scheme.py:187-271, parsing.py:75-111); R = conv
// -------------------------------------------------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(256) k_central(GridDev g, FieldPtrs f, PhysConst c) {
  constexpr int NV = ND + 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z * blockDim.z + threadIdx.z;
  if (i >= g.np[0] || j >= g.np[1] || k >= g.np[2]) return;
  const long long x = g.off + i + (ND > 1 ? j * g.s[1] : 0) + (ND > 2 ? k * g.s[2] : 0);
  double r[NV], q0[NV], u0[ND], div = 0.0;
#pragma unroll
  for (int m = 0; m < NV; m++) { r[m] = 0.0; q0[m] = f.q[m][x]; }
#pragma unroll
  for (int d = 0; d < ND; d++) u0[d] = f.u[d][x];
#pragma unroll
  for (int d = 0; d < ND; d++) {
    const long long s = g.s[d];
    const double um2 = __ldg(f.u[d] + x - 2 * s), um1 = __ldg(f.u[d] + x - s), up1 = __ldg(f.u[d] + x + s), up2 = __ldg(f.u[d] + x + 2 * s);
    div += d1c(um2, um1, up1, up2, c.inv[d]);
#pragma unroll
    for (int m = 0; m < NV; m++) {
      const double qm2 = __ldg(f.q[m] + x - 2 * s), qm1 = __ldg(f.q[m] + x - s), qp1 = __ldg(f.q[m] + x + s), qp2 = __ldg(f.q[m] + x + 2 * s);
      // conservative part d(q_m u_d)/dx_d and advective part u_d dq_m/dx_d
      r[m] += d1c(qm2 * um2, qm1 * um1, qp1 * up1, qp2 * up2, c.inv[d]) + u0[d] * d1c(qm2, qm1, qp1, qp2, c.inv[d]);
    }
    const double pm2 = __ldg(f.p + x - 2 * s), pm1 = __ldg(f.p + x - s), pp1 = __ldg(f.p + x + s), pp2 = __ldg(f.p + x + 2 * s);
    // pressure gradient (momentum d) and pressure work (energy), carried with weight 2 to share the -1/2 below
    r[1 + d] += 2.0 * d1c(pm2, pm1, pp1, pp2, c.inv[d]);
    r[ND + 1] += 2.0 * d1c(pm2 * um2, pm1 * um1, pp1 * up1, pp2 * up2, c.inv[d]);
  }
#pragma unroll
  for (int m = 0; m < NV; m++) f.R[m][x] = -0.5 * (r[m] + q0[m] * div);
}

// -------------------------------------------------------------------------------------------------
// Viscous terms, constant viscosity (StoreSome.py:71-161 / scheme.py:256-327); R += visc
//   tau_ij = (1/Re)(du_i/dx_j + du_j/dx_i - 2/3 delta_ij div u) ; q_j = dT/dx_j /((g-1) Minf^2 Pr Re)
// Mixed derivatives: derivative along the higher direction of the derivative along the lower one.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dev_d1(const double *a, long long x, long long s, double inv) {
  return d1c(__ldg(a + x - 2 * s), __ldg(a + x - s), __ldg(a + x + s), __ldg(a + x + 2 * s), inv);
}
__device__ __forceinline__ double dev_d2(const double *a, long long x, long long s, double inv2) {
  return d2c(__ldg(a + x - 2 * s), __ldg(a + x - s), __ldg(a + x), __ldg(a + x + s), __ldg(a + x + 2 * s), inv2);
}
// d/dx_out ( d a / dx_in )
__device__ __forceinline__ double dev_dmix(const double *a, long long x, long long sin_, double invin, long long sout, double invout) {
  return d1c(dev_d1(a, x - 2 * sout, sin_, invin), dev_d1(a, x - sout, sin_, invin),
             dev_d1(a, x + sout, sin_, invin), dev_d1(a, x + 2 * sout, sin_, invin), invout);
}

template <int ND>
__global__ void __launch_bounds__(256, 2) k_viscous(GridDev g, FieldPtrs f, PhysConst c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z * blockDim.z + threadIdx.z;
  if (i >= g.np[0] || j >= g.np[1] || k >= g.np[2]) return;
  const long long x = g.off + i + (ND > 1 ? j * g.s[1] : 0) + (ND > 2 ? k * g.s[2] : 0);
  const double iRe = 1.0 / c.Re;
  const double kq = iRe * (1.0 / (c.gama - 1.0)) * (1.0 / (c.Minf * c.Minf)) * (1.0 / c.Pr);
  double dv[ND][ND];   // dv[a][b] = d u_a / d x_b
#pragma unroll
  for (int a = 0; a < ND; a++)
#pragma unroll
    for (int b = 0; b < ND; b++) dv[a][b] = dev_d1(f.u[a], x, g.s[b], c.inv[b]);
  double vis[ND], e = 0.0, div = 0.0;
#pragma unroll
  for (int a = 0; a < ND; a++) div += dv[a][a];
#pragma unroll
  for (int a = 0; a < ND; a++) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < ND; b++) {
      if (b == a) s += (4.0 / 3.0) * dev_d2(f.u[a], x, g.s[a], c.inv2[a]);
      else {
        s += dev_d2(f.u[a], x, g.s[b], c.inv2[b]);
        const int in = a < b ? a : b, out = a < b ? b : a;
        s += (1.0 / 3.0) * dev_dmix(f.u[b], x, g.s[in], c.inv[in], g.s[out], c.inv[out]);
      }
    }
    vis[a] = iRe * s;
  }
  double lapT = 0.0;
#pragma unroll
  for (int d = 0; d < ND; d++) lapT += dev_d2(f.T, x, g.s[d], c.inv2[d]);
  e = kq * lapT;
#pragma unroll
  for (int a = 0; a < ND; a++) {
#pragma unroll
    for (int b = a + 1; b < ND; b++) { const double sab = dv[a][b] + dv[b][a]; e += iRe * sab * sab; }
    e += iRe * (2.0 * dv[a][a] - (2.0 / 3.0) * div) * dv[a][a];
    e += vis[a] * __ldg(f.u[a] + x);
  }
  double old[ND + 1];
#pragma unroll
  for (int a = 0; a < ND + 1; a++) old[a] = f.R[1 + a][x];
#pragma unroll
  for (int a = 0; a < ND; a++) f.R[1 + a][x] = old[a] + vis[a];
  f.R[ND + 1][x] = old[ND] + e;
}

// -------------------------------------------------------------------------------------------------
// 3-D viscous terms, shared-memory version (constant viscosity, uniform periodic grid) with the RK update fused in.
// A block owns a 32 x 8 column in (x, y) and marches along z keeping the 5 planes k-2..k+2 of (u0, u1, u2, T) incl. a
// halo of 2 in shared memory (ring buffer), so every stencil value is read from HBM once per column chunk.
// With FUSE_RK the kernel finishes the stage: Residual (flux sweeps) + viscous terms -> low-storage / SBLI RK update of
// q and the RK register; otherwise it leaves Residual += viscous (parity entry point osb_residual).
// -------------------------------------------------------------------------------------------------
// Slab decomposition: the kernel that finishes a stage also delivers the new boundary planes to the neighbour ranks by
// peer stores over NVLink (pointers obtained through CUDA IPC): plane k >= np2-hm goes to the high neighbour's low halo,
// plane k < hp to the low neighbour's high halo (which starts at that neighbour's own slab thickness).
struct PeerPush {
  double *lo[5], *hi[5];     // neighbour's conserved arrays (nullptr: no neighbour on that side)
  int hm, hp;
  int np_lo;                 // slab thickness of the low neighbour (its high halo starts at that plane; slabs may differ by one)
};

// Stream-ordered cross-GPU synchronisation without the host: a rank publishes an epoch number in its neighbours' flag
// words and waits for theirs.  The wait gives up after ~10 s and reports through *err instead of hanging the GPU.
__global__ void k_signal(unsigned long long *a, unsigned long long *b, unsigned long long e) {
  __threadfence_system();
  if (a) *(volatile unsigned long long *)a = e;
  if (b) *(volatile unsigned long long *)b = e;
  __threadfence_system();
}
__global__ void k_wait(const unsigned long long *a, const unsigned long long *b, unsigned long long e, unsigned long long *err) {
  if (*(const volatile unsigned long long *)err) return;      // an earlier wait timed out: do not stall every later stage for 10 s more
  const long long t0 = clock64();
  while ((a && *(const volatile unsigned long long *)a < e) || (b && *(const volatile unsigned long long *)b < e)) {
    __nanosleep(500);
    if (clock64() - t0 > (20LL << 30)) { *err = e; break; }
  }
  __threadfence_system();
}

constexpr int VT_X = 32, VT_Y = 8, VT_HX = VT_X + 4, VT_HY = VT_Y + 4, VT_PLANE = VT_HX * VT_HY, VT_ZC = 32;
constexpr size_t vt_smem_bytes() { return sizeof(double) * 5 * 4 * VT_PLANE; }

// u0, u1, u2, T of one point from its conserved state (constituent relations of the canonical system)
__device__ __forceinline__ void prim_uT(const double *q, const PhysConst &c, double *w) {
  const double irho = rcp_nr(q[0]);
  w[0] = q[1] * irho; w[1] = q[2] * irho; w[2] = q[3] * irho;
  const double p = (c.gama - 1.0) * (q[4] - 0.5 * q[0] * (w[0] * w[0] + w[1] * w[1] + w[2] * w[2]));
  w[3] = c.Minf * c.Minf * c.gama * p * irho;
}

// FROMQ: the planes of (u, T) are computed from q while staging (no constituent-relation kernel, four array reads fewer);
// the new state is then written OUT OF PLACE into the Residual buffers (neighbouring blocks still stage the old q), which
// exchange roles with the q buffers after the launch.
template <int RK, bool FROMQ = false>   // RK: 0 = residual only, 1 = low-storage update (rk_LS.py:139-166), 2 = SBLI update (rk_sbli.py:102-133)
__global__ void __launch_bounds__(VT_X * VT_Y, 2) k_viscous3d_tiled(GridDev g, FieldPtrs f, PhysConst c, double rkA, double rkB, PeerPush pp) {
  extern __shared__ double vt_smem[];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * VT_X + tx;
  // z-chunks are scheduled with the two boundary chunks first: they also perform the peer stores of a decomposed run,
  // which are slower than local stores and would otherwise form the tail of the launch
  const int nzc = gridDim.z, zc = blockIdx.z == 0 ? 0 : (blockIdx.z == 1 ? nzc - 1 : (int)blockIdx.z - 1);
  const int i0 = blockIdx.x * VT_X, j0 = blockIdx.y * VT_Y, k0 = zc * g.zlen;
  const int i = i0 + tx, j = j0 + ty;
  const int kend = min(k0 + g.zlen, g.np[2]);
  const double *src[4] = {f.u[0], f.u[1], f.u[2], f.T};
  auto S = [&](int slot, int v) -> double * { return vt_smem + (size_t)(slot * 4 + v) * VT_PLANE; };
  auto load_plane = [&](int kk) {
    const int slot = (kk + 10) % 5;
    for (int e = tid; e < VT_PLANE; e += VT_X * VT_Y) {
      const int yy = e / VT_HX, xx = e % VT_HX;
      const int gi = i0 + xx - 2, gj = j0 + yy - 2;
      if (gi < g.np[0] + 2 && gj < g.np[1] + 2) {
        const long long x = g.off + gi + gj * g.s[1] + (long long)kk * g.s[2];
        if (FROMQ) {
          double q[5], w[4];
#pragma unroll
          for (int m = 0; m < 5; m++) q[m] = __ldg(f.q[m] + x);
          prim_uT(q, c, w);
#pragma unroll
          for (int v = 0; v < 4; v++) S(slot, v)[e] = w[v];
        } else {
#pragma unroll
          for (int v = 0; v < 4; v++) S(slot, v)[e] = __ldg(src[v] + x);
        }
      }
    }
  };
  for (int kk = k0 - 2; kk <= k0 + 2; kk++) load_plane(kk);
  const double iRe = 1.0 / c.Re;
  const double kq = iRe * (1.0 / (c.gama - 1.0)) * (1.0 / (c.Minf * c.Minf)) * (1.0 / c.Pr);
  const int ce = (ty + 2) * VT_HX + tx + 2;
  constexpr int NPF = (VT_PLANE + VT_X * VT_Y - 1) / (VT_X * VT_Y);
  const bool active = i < g.np[0] && j < g.np[1];
  __syncthreads();
  for (int k = k0; k < kend; k++) {
    // software pipeline: fetch plane k+3 and this point's Residual / RK register / q while plane k is computed
    double pf[NPF][FROMQ ? 5 : 4];
    const bool more = k + 1 < kend;
    if (more) {
#pragma unroll
      for (int it = 0; it < NPF; it++) {
        const int e = tid + it * VT_X * VT_Y;
        const int yy = e / VT_HX, xx = e % VT_HX;
        const int gi = i0 + xx - 2, gj = j0 + yy - 2;
        if (e < VT_PLANE && gi < g.np[0] + 2 && gj < g.np[1] + 2) {
          const long long xg = g.off + gi + gj * g.s[1] + (long long)(k + 3) * g.s[2];
          if (FROMQ) {
#pragma unroll
            for (int m = 0; m < 5; m++) pf[it][m] = __ldg(f.q[m] + xg);
          } else {
#pragma unroll
            for (int v = 0; v < 4; v++) pf[it][v] = __ldg(src[v] + xg);
          }
        }
      }
    }
    const long long x = g.off + i + j * g.s[1] + (long long)k * g.s[2];
    if (more && active && (tx & 15) == 0) {
#pragma unroll
      for (int m = 0; m < 5; m++) {
        prefetch_l2(f.R[m] + x + g.s[2]);
        if (RK != 0) prefetch_l2(f.rk[m] + x + g.s[2]);
        if (RK != 0 && !FROMQ) prefetch_l2(f.q[m] + x + g.s[2]);
      }
    }
    double R[5], q[5], o[5];
    if (active) {
#pragma unroll
      for (int m = 0; m < 5; m++) R[m] = f.R[m][x];
      if (RK != 0) {
#pragma unroll
        for (int m = 0; m < 5; m++) { o[m] = f.rk[m][x]; q[m] = f.q[m][x]; }
      }
    }
    if (active) {
      const double *P[5][4];
#pragma unroll
      for (int dz = 0; dz < 5; dz++)
#pragma unroll
        for (int v = 0; v < 4; v++) P[dz][v] = S((k + dz - 2 + 10) % 5, v) + ce;
      // first derivatives d u_a / d x_b and the Laplacian-like second derivatives at the point
      double dv[3][3], d2[4][3];
#pragma unroll
      for (int v = 0; v < 4; v++) {
        const double *p = P[2][v];
        const double fx = d1c(p[-2], p[-1], p[1], p[2], c.inv[0]);
        const double fy = d1c(p[-2 * VT_HX], p[-VT_HX], p[VT_HX], p[2 * VT_HX], c.inv[1]);
        const double fz = d1c(P[0][v][0], P[1][v][0], P[3][v][0], P[4][v][0], c.inv[2]);
        if (v < 3) { dv[v][0] = fx; dv[v][1] = fy; dv[v][2] = fz; }
        d2[v][0] = d2c(p[-2], p[-1], p[0], p[1], p[2], c.inv2[0]);
        d2[v][1] = d2c(p[-2 * VT_HX], p[-VT_HX], p[0], p[VT_HX], p[2 * VT_HX], c.inv2[1]);
        d2[v][2] = d2c(P[0][v][0], P[1][v][0], p[0], P[3][v][0], P[4][v][0], c.inv2[2]);
      }
      // mixed derivatives: derivative along the higher direction of the derivative along the lower one
      auto dxy = [&](int v) {        // d/dy ( d/dx )
        const double *p = P[2][v];
        double r[4];
        const int oy[4] = {-2 * VT_HX, -VT_HX, VT_HX, 2 * VT_HX};
#pragma unroll
        for (int q = 0; q < 4; q++) r[q] = d1c(p[oy[q] - 2], p[oy[q] - 1], p[oy[q] + 1], p[oy[q] + 2], c.inv[0]);
        return d1c(r[0], r[1], r[2], r[3], c.inv[1]);
      };
      auto dxz = [&](int v) {        // d/dz ( d/dx )
        double r[4];
        const int pz[4] = {0, 1, 3, 4};
#pragma unroll
        for (int q = 0; q < 4; q++) { const double *p = P[pz[q]][v]; r[q] = d1c(p[-2], p[-1], p[1], p[2], c.inv[0]); }
        return d1c(r[0], r[1], r[2], r[3], c.inv[2]);
      };
      auto dyz = [&](int v) {        // d/dz ( d/dy )
        double r[4];
        const int pz[4] = {0, 1, 3, 4};
#pragma unroll
        for (int q = 0; q < 4; q++) { const double *p = P[pz[q]][v]; r[q] = d1c(p[-2 * VT_HX], p[-VT_HX], p[VT_HX], p[2 * VT_HX], c.inv[1]); }
        return d1c(r[0], r[1], r[2], r[3], c.inv[2]);
      };
      double vis[3];
      vis[0] = iRe * ((4.0 / 3.0) * d2[0][0] + d2[0][1] + d2[0][2] + (1.0 / 3.0) * (dxy(1) + dxz(2)));
      vis[1] = iRe * (d2[1][0] + (4.0 / 3.0) * d2[1][1] + d2[1][2] + (1.0 / 3.0) * (dxy(0) + dyz(2)));
      vis[2] = iRe * (d2[2][0] + d2[2][1] + (4.0 / 3.0) * d2[2][2] + (1.0 / 3.0) * (dxz(0) + dyz(1)));
      const double div = dv[0][0] + dv[1][1] + dv[2][2];
      double e = kq * (d2[3][0] + d2[3][1] + d2[3][2]);
#pragma unroll
      for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = a + 1; b < 3; b++) { const double sab = dv[a][b] + dv[b][a]; e += iRe * sab * sab; }
        e += iRe * (2.0 * dv[a][a] - (2.0 / 3.0) * div) * dv[a][a];
        e += vis[a] * P[2][a][0];
      }
      R[1] += vis[0]; R[2] += vis[1]; R[3] += vis[2]; R[4] += e;
      if (RK == 0) {
#pragma unroll
        for (int m = 1; m < 5; m++) f.R[m][x] = R[m];
      } else {
#pragma unroll
        for (int m = 0; m < 5; m++) {
          double qn;
          if (RK == 1) { const double t = c.dt * R[m] + rkA * o[m]; f.rk[m][x] = t; qn = rkB * t + q[m]; }
          else { qn = c.dt * rkB * R[m] + o[m]; f.rk[m][x] = c.dt * rkA * R[m] + o[m]; }
          if (FROMQ) f.R[m][x] = qn; else f.q[m][x] = qn;
          // fused halo exchange: peer stores of the new boundary planes
          if (pp.hi[m] && k >= g.np[2] - pp.hm) pp.hi[m][x - (long long)g.np[2] * g.s[2]] = qn;
          if (pp.lo[m] && k < pp.hp) pp.lo[m][x + (long long)pp.np_lo * g.s[2]] = qn;
        }
      }
    }
    __syncthreads();                 // every thread is done reading the slot of plane k-2
    if (more) {
      const int slot = (k + 3 + 10) % 5;
#pragma unroll
      for (int it = 0; it < NPF; it++) {
        const int e = tid + it * VT_X * VT_Y;
        const int yy = e / VT_HX, xx = e % VT_HX;
        const int gi = i0 + xx - 2, gj = j0 + yy - 2;
        if (e < VT_PLANE && gi < g.np[0] + 2 && gj < g.np[1] + 2) {
          if (FROMQ) {
            double w[4];
            prim_uT(pf[it], c, w);
#pragma unroll
            for (int v = 0; v < 4; v++) S(slot, v)[e] = w[v];
          } else {
#pragma unroll
            for (int v = 0; v < 4; v++) S(slot, v)[e] = pf[it][v];
          }
        }
      }
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
// The same stage kernel (constituent relations + viscous terms + RK update out of place + peer stores) with a TMA + mbarrier
// plane pipeline: the raw conserved variables of plane k+3 / k+4 (tile + halo 2, one box of 38 x 12 per array) travel into a
// two-stage ring by cp.async.bulk.tensor while plane k is computed; one elected thread issues, everybody waits on the
// mbarrier of the stage, converts its share of the landed plane to (u0, u1, u2, T) and moves on.  The per-thread staging
// loads, their address arithmetic and the 20 registers that carried them are gone; this point's Residual / RK register / q
// lines of the NEXT plane are requested into L2 one iteration ahead.  Needs an even padded x-extent (TMA strides).
// -------------------------------------------------------------------------------------------------
constexpr int VTM_BX = 38, VTM_RAW = 464;          // box width (lanes -3 .. 34: 16-byte aligned start), doubles per landed array (3648 B padded to 29 x 128 B)
constexpr size_t vtm_smem_bytes() { return sizeof(double) * (2 * 5 * VTM_RAW + 5 * 4 * VT_PLANE) + 64; }
__device__ __forceinline__ double ldg_f64_volatile(const double *p) {      // not sunk to its use by the compiler
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

template <int RK>   // 1 = low-storage update (rk_LS.py:139-166), 2 = SBLI update (rk_sbli.py:102-133)
__global__ void __launch_bounds__(VT_X * VT_Y, 2) k_viscous3d_tma(GridDev g, FieldPtrs f, PhysConst c, double rkA, double rkB, PeerPush pp,
                                                                 const __grid_constant__ TmaMaps5 maps) {
  extern __shared__ __align__(128) double vtm_smem[];
  double *raw = vtm_smem;                                   // [2][5][VTM_RAW]
  double *conv = raw + 2 * 5 * VTM_RAW;                     // [5 slots][4][VT_PLANE]
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(conv + 5 * 4 * VT_PLANE);   // [2]
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * VT_X + tx;
  const int nzc = gridDim.z, zc = blockIdx.z == 0 ? 0 : (blockIdx.z == 1 ? nzc - 1 : (int)blockIdx.z - 1);   // boundary chunks first (peer stores)
  const int i0 = blockIdx.x * VT_X, j0 = blockIdx.y * VT_Y, k0 = zc * g.zlen;
  const int i = i0 + tx, j = j0 + ty;
  const int kend = min(k0 + g.zlen, g.np[2]);
  const int NP = (kend - k0) + 4;                           // planes k0-2 .. kend+1, numbered n = 0 .. NP-1
  if (tid == 0) {
    mbar_init(bar, 1); mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int n) {                                 // thread 0: plane n -> raw stage n & 1
    unsigned long long *b = bar + (n & 1);
    mbar_expect_tx(b, 5 * VTM_BX * VT_HY * (unsigned)sizeof(double));
#pragma unroll
    for (int m = 0; m < 5; m++)
      tma_load_3d(raw + ((n & 1) * 5 + m) * VTM_RAW, maps.m[m], b, i0 - 3 + g.h, j0 - 2 + g.h, k0 - 2 + n + g.h);
  };
  auto convert = [&](int n) {                               // landed plane n -> (u0, u1, u2, T) in ring slot n % 5
    const double *rw = raw + (n & 1) * 5 * VTM_RAW;
    double *cv = conv + (n % 5) * 4 * VT_PLANE;
    for (int e = tid; e < VT_PLANE; e += VT_X * VT_Y) {
      const int yy = e / VT_HX, xx = e - yy * VT_HX;
      const int r = yy * VTM_BX + xx + 1;
      double q[5], w[4];
#pragma unroll
      for (int m = 0; m < 5; m++) q[m] = rw[m * VTM_RAW + r];
      prim_uT(q, c, w);
#pragma unroll
      for (int v = 0; v < 4; v++) cv[v * VT_PLANE + e] = w[v];
    }
  };
  if (tid == 0) { issue(0); if (NP > 1) issue(1); }
  for (int n = 0; n < 5; n++) {                             // prologue: planes k0-2 .. k0+2
    mbar_wait(bar + (n & 1), (n >> 1) & 1);
    convert(n);
    __syncthreads();
    if (tid == 0 && n + 2 < NP) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(n + 2); }
  }
  const double iRe = 1.0 / c.Re;
  const double kq = iRe * (1.0 / (c.gama - 1.0)) * (1.0 / (c.Minf * c.Minf)) * (1.0 / c.Pr);
  const int ce = (ty + 2) * VT_HX + tx + 2;
  const bool active = i < g.np[0] && j < g.np[1];
  const long long xcol = g.off + i + j * g.s[1];
  if (active && (tx & 15) == 0) {                           // this point's Residual / RK register lines of the first plane into L2
    const long long x = xcol + (long long)k0 * g.s[2];
#pragma unroll
    for (int m = 0; m < 5; m++) { prefetch_l2(f.R[m] + x); prefetch_l2(f.rk[m] + x); }
  }
  int s0 = 0;                                               // ring slot of plane k-2 (planes k-2 .. k+2 in slots s0, s0+1, .. mod 5)
  // z-derivatives of x- and y-derivatives: this point's d/dx u0, d/dx u2, d/dy u1, d/dy u2 of the planes k-2 .. k+2 ride along in
  // registers (one new plane per iteration), so that d/dz(d/dx .) and d/dz(d/dy .) cost no shared-memory loads (the kernel is
  // bound by the shared-memory pipe: ncu r02, short_scoreboard 4.2 and mio_throttle 2.9 per issue, wavefronts 65 % of peak)
  double gx0[5], gx2[5], gy1[5], gy2[5];
  auto plane_grads = [&](int slot, double &ax0, double &ax2, double &ay1, double &ay2) {
    const double *p0 = conv + (slot * 4 + 0) * VT_PLANE + ce, *p1 = p0 + VT_PLANE, *p2 = p1 + VT_PLANE;
    ax0 = d1c(p0[-2], p0[-1], p0[1], p0[2], c.inv[0]);
    ax2 = d1c(p2[-2], p2[-1], p2[1], p2[2], c.inv[0]);
    ay1 = d1c(p1[-2 * VT_HX], p1[-VT_HX], p1[VT_HX], p1[2 * VT_HX], c.inv[1]);
    ay2 = d1c(p2[-2 * VT_HX], p2[-VT_HX], p2[VT_HX], p2[2 * VT_HX], c.inv[1]);
  };
  if (active) {
#pragma unroll
    for (int n = 0; n < 4; n++) plane_grads(n, gx0[n + 1], gx2[n + 1], gy1[n + 1], gy2[n + 1]);   // shifted down at the top of the first iteration
  }
#pragma unroll 1
  for (int k = k0; k < kend; k++) {
    const long long x = xcol + (long long)k * g.s[2];
    const bool more = k + 1 < kend;
    // this point's Residual, RK register and q: requested at the top of the iteration (volatile: the compiler would sink the
    // loads to their use at the end and expose the latency), the lines were brought into L2 one iteration earlier
    if (active && more && (tx & 15) == 0) {
#pragma unroll
      for (int m = 0; m < 5; m++) { prefetch_l2(f.R[m] + x + g.s[2]); prefetch_l2(f.rk[m] + x + g.s[2]); }
    }
    if (active) {
#pragma unroll
      for (int n = 0; n < 4; n++) { gx0[n] = gx0[n + 1]; gx2[n] = gx2[n + 1]; gy1[n] = gy1[n + 1]; gy2[n] = gy2[n + 1]; }
      { int sl = s0 + 4; sl = sl >= 5 ? sl - 5 : sl; plane_grads(sl, gx0[4], gx2[4], gy1[4], gy2[4]); }
      const double *P[5][4];
#pragma unroll
      for (int dz = 0; dz < 5; dz++) {
        int sl = s0 + dz; sl = sl >= 5 ? sl - 5 : sl;
#pragma unroll
        for (int v = 0; v < 4; v++) P[dz][v] = conv + (sl * 4 + v) * VT_PLANE + ce;
      }
      double dv[3][3], d2[4][3];
#pragma unroll
      for (int v = 0; v < 4; v++) {
        const double *p = P[2][v];
        const double fx = d1c(p[-2], p[-1], p[1], p[2], c.inv[0]);
        const double fy = d1c(p[-2 * VT_HX], p[-VT_HX], p[VT_HX], p[2 * VT_HX], c.inv[1]);
        const double fz = d1c(P[0][v][0], P[1][v][0], P[3][v][0], P[4][v][0], c.inv[2]);
        if (v < 3) { dv[v][0] = fx; dv[v][1] = fy; dv[v][2] = fz; }
        d2[v][0] = d2c(p[-2], p[-1], p[0], p[1], p[2], c.inv2[0]);
        d2[v][1] = d2c(p[-2 * VT_HX], p[-VT_HX], p[0], p[VT_HX], p[2 * VT_HX], c.inv2[1]);
        d2[v][2] = d2c(P[0][v][0], P[1][v][0], p[0], P[3][v][0], P[4][v][0], c.inv2[2]);
      }
      auto dxy = [&](int v) {        // d/dy ( d/dx )
        const double *p = P[2][v];
        double r[4];
        const int oy[4] = {-2 * VT_HX, -VT_HX, VT_HX, 2 * VT_HX};
#pragma unroll
        for (int q = 0; q < 4; q++) r[q] = d1c(p[oy[q] - 2], p[oy[q] - 1], p[oy[q] + 1], p[oy[q] + 2], c.inv[0]);
        return d1c(r[0], r[1], r[2], r[3], c.inv[1]);
      };
      // d/dz(d/dx u2), d/dz(d/dx u0), d/dz(d/dy u2), d/dz(d/dy u1) from the register rings
      const double dxz2 = d1c(gx2[0], gx2[1], gx2[3], gx2[4], c.inv[2]), dxz0 = d1c(gx0[0], gx0[1], gx0[3], gx0[4], c.inv[2]);
      const double dyz2 = d1c(gy2[0], gy2[1], gy2[3], gy2[4], c.inv[2]), dyz1 = d1c(gy1[0], gy1[1], gy1[3], gy1[4], c.inv[2]);
      double vis[3];
      vis[0] = iRe * ((4.0 / 3.0) * d2[0][0] + d2[0][1] + d2[0][2] + (1.0 / 3.0) * (dxy(1) + dxz2));
      vis[1] = iRe * (d2[1][0] + (4.0 / 3.0) * d2[1][1] + d2[1][2] + (1.0 / 3.0) * (dxy(0) + dyz2));
      vis[2] = iRe * (d2[2][0] + d2[2][1] + (4.0 / 3.0) * d2[2][2] + (1.0 / 3.0) * (dxz0 + dyz1));
      const double div = dv[0][0] + dv[1][1] + dv[2][2];
      double e = kq * (d2[3][0] + d2[3][1] + d2[3][2]);
#pragma unroll
      for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = a + 1; b < 3; b++) { const double sab = dv[a][b] + dv[b][a]; e += iRe * sab * sab; }
        e += iRe * (2.0 * dv[a][a] - (2.0 / 3.0) * div) * dv[a][a];
        e += vis[a] * P[2][a][0];
      }
      // this point's Residual, RK register and q: the lines were brought into L2 one iteration earlier (q: by the plane's box);
      // loaded only now -- held across the derivative evaluation they cost 30 registers and spill
      double cR[5], cO[5], cQ[5];
#pragma unroll
      for (int m = 0; m < 5; m++) { cR[m] = ldg_f64_volatile(f.R[m] + x); cO[m] = ldg_f64_volatile(f.rk[m] + x); cQ[m] = ldg_f64_volatile(f.q[m] + x); }
      cR[1] += vis[0]; cR[2] += vis[1]; cR[3] += vis[2]; cR[4] += e;
#pragma unroll
      for (int m = 0; m < 5; m++) {
        double qn;
        if (RK == 1) { const double t = c.dt * cR[m] + rkA * cO[m]; f.rk[m][x] = t; qn = rkB * t + cQ[m]; }
        else { qn = c.dt * rkB * cR[m] + cO[m]; f.rk[m][x] = c.dt * rkA * cR[m] + cO[m]; }
        f.R[m][x] = qn;                                      // out of place: the Residual buffers become the new q
        if (pp.hi[m] && k >= g.np[2] - pp.hm) pp.hi[m][x - (long long)g.np[2] * g.s[2]] = qn;
        if (pp.lo[m] && k < pp.hp) pp.lo[m][x + (long long)pp.np_lo * g.s[2]] = qn;
      }
    }
    __syncthreads();                                        // every thread is done reading the slot of plane k-2
    if (more) {
      const int n = k - k0 + 5;                             // plane k+3
      mbar_wait(bar + (n & 1), (n >> 1) & 1);
      convert(n);
      __syncthreads();
      if (tid == 0 && n + 2 < NP) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(n + 2); }
      s0 = s0 == 4 ? 0 : s0 + 1;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Central(4) Navier-Stokes stage in ONE kernel (3-D, constant viscosity, uniform grid; BASELINE config 2):
// constituent relations + skew-symmetric convective terms + viscous terms + RK update.  A block marches along z over a
// 32 x 8 column with the planes k-2..k+2 of (rho, E, u0, u1, u2, p, T) in shared memory (computed from q while staging),
// so per stage q is read once (plus tile halos) and q, RK register are written once.  The update is out of place
// (q_in -> q_out): neighbouring blocks still read the old values of the points this block updates.
// -------------------------------------------------------------------------------------------------
constexpr int CT_NV = 7, CT_RHO = 0, CT_E = 1, CT_U = 2, CT_P = 5, CT_T = 6;
constexpr size_t ct_smem_bytes(bool tma = false) { return sizeof(double) * (5 * CT_NV * VT_PLANE + (tma ? 2 * 5 * 464 : 0)) + 64; }
struct QPtrs { double *q[5]; };

// TMA = true: the raw conserved planes arrive through the TMA + mbarrier pipeline of k_viscous3d_tma (two-stage ring of
// 38 x 12 boxes per array) instead of per-thread loads, the RK register is loaded just before the update and the
// z-derivatives of x/y-derivatives come from register rings; needs an even padded x-extent.
template <int RK, bool PUSH, bool TMA>   // RK 1 = low-storage update, 2 = SBLI update (first_stage folds the "Save equations" loop: old = q); PUSH: slab run
__global__ void __launch_bounds__(VT_X * VT_Y, 1) k_central3d_fused(GridDev g, QPtrs qin, QPtrs qout, QPtrs rkreg, PhysConst c,
                                                                   double rkA, double rkB, int first_stage, PeerPush pp,
                                                                   const __grid_constant__ TmaMaps5 maps) {
  extern __shared__ __align__(128) double ct_smem_all[];
  double *raw = ct_smem_all;                                              // TMA: [2][5][VTM_RAW] landing zone
  double *ct_smem = ct_smem_all + (TMA ? 2 * 5 * VTM_RAW : 0);           // [5 slots][CT_NV][VT_PLANE]
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(ct_smem + 5 * CT_NV * VT_PLANE);
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * VT_X + tx;
  const int i0 = blockIdx.x * VT_X, j0 = blockIdx.y * VT_Y, k0 = blockIdx.z * g.zlen;
  const int i = i0 + tx, j = j0 + ty;
  const int kend = min(k0 + g.zlen, g.np[2]);
  constexpr int NPF = (VT_PLANE + VT_X * VT_Y - 1) / (VT_X * VT_Y);
  auto S = [&](int slot, int v) -> double * { return ct_smem + (size_t)(slot * CT_NV + v) * VT_PLANE; };
  auto in_tile = [&](int e, long long &xg, int kk) -> bool {
    const int yy = e / VT_HX, xx = e % VT_HX;
    const int gi = i0 + xx - 2, gj = j0 + yy - 2;
    xg = g.off + gi + gj * g.s[1] + (long long)kk * g.s[2];
    return e < VT_PLANE && gi < g.np[0] + 2 && gj < g.np[1] + 2;
  };
  auto stage = [&](int slot, int e, const double *q) {     // constituent relations of one staged point
    const double rho = q[0], irho = TMA ? rcp_nr(rho) : 1.0 / rho;
    const double u0 = q[1] * irho, u1 = q[2] * irho, u2 = q[3] * irho;
    const double p = (c.gama - 1.0) * (q[4] - 0.5 * rho * (u0 * u0 + u1 * u1 + u2 * u2));
    S(slot, CT_RHO)[e] = rho; S(slot, CT_E)[e] = q[4];
    S(slot, CT_U)[e] = u0; S(slot, CT_U + 1)[e] = u1; S(slot, CT_U + 2)[e] = u2;
    S(slot, CT_P)[e] = p; S(slot, CT_T)[e] = c.Minf * c.Minf * c.gama * p * irho;
  };
  const int NP = (kend - k0) + 4;                           // TMA: planes k0-2 .. kend+1, numbered n = 0 .. NP-1
  auto issue = [&](int n) {                                 // thread 0: plane n -> raw stage n & 1
    unsigned long long *b = bar + (n & 1);
    mbar_expect_tx(b, 5 * VTM_BX * VT_HY * (unsigned)sizeof(double));
#pragma unroll
    for (int m = 0; m < 5; m++)
      tma_load_3d(raw + ((n & 1) * 5 + m) * VTM_RAW, maps.m[m], b, i0 - 3 + g.h, j0 - 2 + g.h, k0 - 2 + n + g.h);
  };
  auto convert = [&](int n) {                               // landed plane n -> staged values in ring slot of that plane
    const double *rw = raw + (n & 1) * 5 * VTM_RAW;
    for (int e = tid; e < VT_PLANE; e += VT_X * VT_Y) {
      const int yy = e / VT_HX, xx = e - yy * VT_HX;
      double q[5];
#pragma unroll
      for (int m = 0; m < 5; m++) q[m] = rw[m * VTM_RAW + yy * VTM_BX + xx + 1];
      stage((k0 - 2 + n + 10) % 5, e, q);
    }
  };
  if (TMA) {
    if (tid == 0) {
      mbar_init(bar, 1); mbar_init(bar + 1, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) { issue(0); issue(1); }
    for (int n = 0; n < 5; n++) {
      mbar_wait(bar + (n & 1), (n >> 1) & 1);
      convert(n);
      __syncthreads();
      if (tid == 0 && n + 2 < NP) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(n + 2); }
    }
  } else {
    for (int kk = k0 - 2; kk <= k0 + 2; kk++)
      for (int it = 0; it < NPF; it++) {
        long long xg;
        const int e = tid + it * VT_X * VT_Y;
        if (in_tile(e, xg, kk)) {
          double q[5];
#pragma unroll
          for (int m = 0; m < 5; m++) q[m] = __ldg(qin.q[m] + xg);
          stage((kk + 10) % 5, e, q);
        }
      }
  }
  const double iRe = 1.0 / c.Re;
  const double kq = iRe * (1.0 / (c.gama - 1.0)) * (1.0 / (c.Minf * c.Minf)) * (1.0 / c.Pr);
  const int ce = (ty + 2) * VT_HX + tx + 2;
  const bool active = i < g.np[0] && j < g.np[1];
  __syncthreads();
  // TMA: this point's d/dx u0, d/dx u2, d/dy u1, d/dy u2 of the planes k-2 .. k+2 in registers (see k_viscous3d_tma)
  double gx0[5], gx2[5], gy1[5], gy2[5];
  auto plane_grads = [&](int kk, double &ax0, double &ax2, double &ay1, double &ay2) {
    const double *p0 = S((kk + 10) % 5, CT_U) + ce, *p1 = p0 + VT_PLANE, *p2 = p1 + VT_PLANE;
    ax0 = d1c(p0[-2], p0[-1], p0[1], p0[2], c.inv[0]);
    ax2 = d1c(p2[-2], p2[-1], p2[1], p2[2], c.inv[0]);
    ay1 = d1c(p1[-2 * VT_HX], p1[-VT_HX], p1[VT_HX], p1[2 * VT_HX], c.inv[1]);
    ay2 = d1c(p2[-2 * VT_HX], p2[-VT_HX], p2[VT_HX], p2[2 * VT_HX], c.inv[1]);
  };
  if (TMA && active) {
#pragma unroll
    for (int n = 0; n < 4; n++) plane_grads(k0 - 2 + n, gx0[n + 1], gx2[n + 1], gy1[n + 1], gy2[n + 1]);
  }
#pragma unroll 1
  for (int k = k0; k < kend; k++) {
    // software pipeline: raw q of plane k+3 and this point's RK register travel while plane k is computed
    double pf[NPF][5];
    const bool more = k + 1 < kend;
    if (!TMA && more) {
#pragma unroll
      for (int it = 0; it < NPF; it++) {
        long long xg;
        if (in_tile(tid + it * VT_X * VT_Y, xg, k + 3)) {
#pragma unroll
          for (int m = 0; m < 5; m++) pf[it][m] = __ldg(qin.q[m] + xg);
        }
      }
    }
    const long long x = g.off + i + j * g.s[1] + (long long)k * g.s[2];
    if (more && active && (tx & 15) == 0 && !(RK == 2 && first_stage)) {
#pragma unroll
      for (int m = 0; m < 5; m++) prefetch_l2(rkreg.q[m] + x + g.s[2]);
    }
    double o[5];
    if (!TMA && active && !(RK == 2 && first_stage)) {
#pragma unroll
      for (int m = 0; m < 5; m++) o[m] = rkreg.q[m][x];
    }
    if (active) {
      if (TMA) {
#pragma unroll
        for (int n = 0; n < 4; n++) { gx0[n] = gx0[n + 1]; gx2[n] = gx2[n + 1]; gy1[n] = gy1[n + 1]; gy2[n] = gy2[n + 1]; }
        plane_grads(k + 2, gx0[4], gx2[4], gy1[4], gy2[4]);
      }
      const double *P[5][CT_NV];
#pragma unroll
      for (int dz = 0; dz < 5; dz++)
#pragma unroll
        for (int v = 0; v < CT_NV; v++) P[dz][v] = S((k + dz - 2 + 10) % 5, v) + ce;
      const double rho0 = P[2][CT_RHO][0], E0 = P[2][CT_E][0];
      const double uc[3] = {P[2][CT_U][0], P[2][CT_U + 1][0], P[2][CT_U + 2][0]};
      const double qc[5] = {rho0, rho0 * uc[0], rho0 * uc[1], rho0 * uc[2], E0};
      // ---- skew-symmetric convective terms (parsing.py:75-111, scheme.py:187-271)
      double cons[5] = {0, 0, 0, 0, 0}, adv[5] = {0, 0, 0, 0, 0}, div = 0.0, dp[3], dpu = 0.0;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        double qn[4][5], un[4], pn[4];
#pragma unroll
        for (int n = 0; n < 4; n++) {
          const int s = n < 2 ? n - 2 : n - 1;                    // -2, -1, +1, +2
          const int off = d == 0 ? s : (d == 1 ? s * VT_HX : 0);
          const int dz = d == 2 ? 2 + s : 2;
          const double r = P[dz][CT_RHO][off];
          const double v0 = P[dz][CT_U][off], v1 = P[dz][CT_U + 1][off], v2 = P[dz][CT_U + 2][off];
          un[n] = d == 0 ? v0 : (d == 1 ? v1 : v2);
          pn[n] = P[dz][CT_P][off];
          qn[n][0] = r; qn[n][1] = r * v0; qn[n][2] = r * v1; qn[n][3] = r * v2; qn[n][4] = P[dz][CT_E][off];
        }
        div += d1c(un[0], un[1], un[2], un[3], c.inv[d]);
#pragma unroll
        for (int m = 0; m < 5; m++) {
          cons[m] += d1c(qn[0][m] * un[0], qn[1][m] * un[1], qn[2][m] * un[2], qn[3][m] * un[3], c.inv[d]);
          adv[m] += uc[d] * d1c(qn[0][m], qn[1][m], qn[2][m], qn[3][m], c.inv[d]);
        }
        dp[d] = d1c(pn[0], pn[1], pn[2], pn[3], c.inv[d]);
        dpu += d1c(pn[0] * un[0], pn[1] * un[1], pn[2] * un[2], pn[3] * un[3], c.inv[d]);
      }
      double R[5];
#pragma unroll
      for (int m = 0; m < 5; m++) R[m] = -0.5 * (cons[m] + adv[m] + qc[m] * div);
      R[1] -= dp[0]; R[2] -= dp[1]; R[3] -= dp[2]; R[4] -= dpu;
      // ---- viscous terms (same arithmetic as k_viscous3d_tiled)
      const int VI[4] = {CT_U, CT_U + 1, CT_U + 2, CT_T};
      double dv[3][3], d2[4][3];
#pragma unroll
      for (int v = 0; v < 4; v++) {
        const double *p = P[2][VI[v]];
        const double fx = d1c(p[-2], p[-1], p[1], p[2], c.inv[0]);
        const double fy = d1c(p[-2 * VT_HX], p[-VT_HX], p[VT_HX], p[2 * VT_HX], c.inv[1]);
        const double fz = d1c(P[0][VI[v]][0], P[1][VI[v]][0], P[3][VI[v]][0], P[4][VI[v]][0], c.inv[2]);
        if (v < 3) { dv[v][0] = fx; dv[v][1] = fy; dv[v][2] = fz; }
        d2[v][0] = d2c(p[-2], p[-1], p[0], p[1], p[2], c.inv2[0]);
        d2[v][1] = d2c(p[-2 * VT_HX], p[-VT_HX], p[0], p[VT_HX], p[2 * VT_HX], c.inv2[1]);
        d2[v][2] = d2c(P[0][VI[v]][0], P[1][VI[v]][0], p[0], P[3][VI[v]][0], P[4][VI[v]][0], c.inv2[2]);
      }
      auto dxy = [&](int v) {
        const double *p = P[2][VI[v]];
        double r[4];
        const int oy[4] = {-2 * VT_HX, -VT_HX, VT_HX, 2 * VT_HX};
#pragma unroll
        for (int q = 0; q < 4; q++) r[q] = d1c(p[oy[q] - 2], p[oy[q] - 1], p[oy[q] + 1], p[oy[q] + 2], c.inv[0]);
        return d1c(r[0], r[1], r[2], r[3], c.inv[1]);
      };
      auto dxz = [&](int v) {
        double r[4];
        const int pz[4] = {0, 1, 3, 4};
#pragma unroll
        for (int q = 0; q < 4; q++) { const double *p = P[pz[q]][VI[v]]; r[q] = d1c(p[-2], p[-1], p[1], p[2], c.inv[0]); }
        return d1c(r[0], r[1], r[2], r[3], c.inv[2]);
      };
      auto dyz = [&](int v) {
        double r[4];
        const int pz[4] = {0, 1, 3, 4};
#pragma unroll
        for (int q = 0; q < 4; q++) { const double *p = P[pz[q]][VI[v]]; r[q] = d1c(p[-2 * VT_HX], p[-VT_HX], p[VT_HX], p[2 * VT_HX], c.inv[1]); }
        return d1c(r[0], r[1], r[2], r[3], c.inv[2]);
      };
      double vis[3];
      if (TMA) {     // d/dz(d/dx .), d/dz(d/dy .) from the register rings
        const double dxz2 = d1c(gx2[0], gx2[1], gx2[3], gx2[4], c.inv[2]), dxz0 = d1c(gx0[0], gx0[1], gx0[3], gx0[4], c.inv[2]);
        const double dyz2 = d1c(gy2[0], gy2[1], gy2[3], gy2[4], c.inv[2]), dyz1 = d1c(gy1[0], gy1[1], gy1[3], gy1[4], c.inv[2]);
        vis[0] = iRe * ((4.0 / 3.0) * d2[0][0] + d2[0][1] + d2[0][2] + (1.0 / 3.0) * (dxy(1) + dxz2));
        vis[1] = iRe * (d2[1][0] + (4.0 / 3.0) * d2[1][1] + d2[1][2] + (1.0 / 3.0) * (dxy(0) + dyz2));
        vis[2] = iRe * (d2[2][0] + d2[2][1] + (4.0 / 3.0) * d2[2][2] + (1.0 / 3.0) * (dxz0 + dyz1));
      } else {
        vis[0] = iRe * ((4.0 / 3.0) * d2[0][0] + d2[0][1] + d2[0][2] + (1.0 / 3.0) * (dxy(1) + dxz(2)));
        vis[1] = iRe * (d2[1][0] + (4.0 / 3.0) * d2[1][1] + d2[1][2] + (1.0 / 3.0) * (dxy(0) + dyz(2)));
        vis[2] = iRe * (d2[2][0] + d2[2][1] + (4.0 / 3.0) * d2[2][2] + (1.0 / 3.0) * (dxz(0) + dyz(1)));
      }
      const double dvg = dv[0][0] + dv[1][1] + dv[2][2];
      double e = kq * (d2[3][0] + d2[3][1] + d2[3][2]);
#pragma unroll
      for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = a + 1; b < 3; b++) { const double sab = dv[a][b] + dv[b][a]; e += iRe * sab * sab; }
        e += iRe * (2.0 * dv[a][a] - (2.0 / 3.0) * dvg) * dv[a][a];
        e += vis[a] * uc[a];
      }
      R[1] += vis[0]; R[2] += vis[1]; R[3] += vis[2]; R[4] += e;
      // ---- RK update, out of place
      if (TMA && !(RK == 2 && first_stage)) {       // RK register: line prefetched into L2 one plane earlier, loaded only now
#pragma unroll
        for (int m = 0; m < 5; m++) o[m] = ldg_f64_volatile(rkreg.q[m] + x);
      }
#pragma unroll
      for (int m = 0; m < 5; m++) {
        double qn;
        if (RK == 1) { const double t = c.dt * R[m] + rkA * o[m]; rkreg.q[m][x] = t; qn = rkB * t + qc[m]; }
        else { const double old = first_stage ? qc[m] : o[m]; qn = c.dt * rkB * R[m] + old; rkreg.q[m][x] = c.dt * rkA * R[m] + old; }
        qout.q[m][x] = qn;
        // slab decomposition: new boundary planes go straight into the neighbours' (next) q buffers
        if (PUSH) {
          if (pp.hi[m] && k >= g.np[2] - pp.hm) pp.hi[m][x - (long long)g.np[2] * g.s[2]] = qn;
          if (pp.lo[m] && k < pp.hp) pp.lo[m][x + (long long)pp.np_lo * g.s[2]] = qn;
        }
      }
    }
    __syncthreads();
    if (TMA) {
      if (more) {
        const int n = k - k0 + 5;                           // plane k+3
        mbar_wait(bar + (n & 1), (n >> 1) & 1);
        convert(n);
        __syncthreads();
        if (tid == 0 && n + 2 < NP) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(n + 2); }
      }
      continue;
    }
    if (more) {
#pragma unroll
      for (int it = 0; it < NPF; it++) {
        long long xg;
        const int e = tid + it * VT_X * VT_Y;
        if (in_tile(e, xg, k + 3)) stage((k + 3 + 10) % 5, e, pf[it]);
      }
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
// Runge-Kutta updates (rk_sbli.py:102-133, rk_LS.py:139-166)
// -------------------------------------------------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(256) k_rk_ls(GridDev g, FieldPtrs f, double dt, double A, double B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= g.np[0] || j >= g.np[1] || k >= g.np[2]) return;
  const long long x = g.off + i + (ND > 1 ? j * g.s[1] : 0) + (ND > 2 ? k * g.s[2] : 0);
  double r[ND + 2], k_[ND + 2], q[ND + 2];
#pragma unroll
  for (int m = 0; m < ND + 2; m++) { r[m] = __ldg(f.R[m] + x); k_[m] = f.rk[m][x]; q[m] = f.q[m][x]; }
#pragma unroll
  for (int m = 0; m < ND + 2; m++) {
    const double t = dt * r[m] + A * k_[m];
    f.rk[m][x] = t;
    f.q[m][x] = B * t + q[m];
  }
}
template <int ND>
__global__ void __launch_bounds__(256) k_rk_sbli(GridDev g, FieldPtrs f, double dt, double rkold, double rknew) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= g.np[0] || j >= g.np[1] || k >= g.np[2]) return;
  const long long x = g.off + i + (ND > 1 ? j * g.s[1] : 0) + (ND > 2 ? k * g.s[2] : 0);
  double r[ND + 2], o[ND + 2];
#pragma unroll
  for (int m = 0; m < ND + 2; m++) { r[m] = __ldg(f.R[m] + x); o[m] = f.rk[m][x]; }
#pragma unroll
  for (int m = 0; m < ND + 2; m++) {
    f.q[m][x] = dt * rknew * r[m] + o[m];
    f.rk[m][x] = dt * rkold * r[m] + o[m];
  }
}
template <int ND>
__global__ void __launch_bounds__(256) k_rk_save(GridDev g, FieldPtrs f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= g.np[0] || j >= g.np[1] || k >= g.np[2]) return;
  const long long x = g.off + i + (ND > 1 ? j * g.s[1] : 0) + (ND > 2 ? k * g.s[2] : 0);
  double q[ND + 2];
#pragma unroll
  for (int m = 0; m < ND + 2; m++) q[m] = f.q[m][x];
#pragma unroll
  for (int m = 0; m < ND + 2; m++) f.rk[m][x] = q[m];
}

// -------------------------------------------------------------------------------------------------
// General path: one-sided closures, stretched-grid metrics, variable viscosity (BASELINE config 4)
// -------------------------------------------------------------------------------------------------
// first / second xi-derivative along direction dir at linear index x, grid index idx of n points
__device__ __forceinline__ double gd1(const Closures &cl, const double *a, long long x, long long s, double inv, int dir, int idx, int n) {
  if (cl.on[dir][0] && idx < cl.nr1) {
    double r = 0.0; const long long x0 = x - idx * s;
    for (int p = 0; p < cl.np1; p++) r += cl.d1[idx * cl.np1 + p] * __ldg(a + x0 + p * s);
    return inv * r;
  }
  if (cl.on[dir][1] && n - 1 - idx < cl.nr1) {
    const int row = n - 1 - idx; double r = 0.0; const long long x0 = x + row * s;
    for (int p = 0; p < cl.np1; p++) r -= cl.d1[row * cl.np1 + p] * __ldg(a + x0 - p * s);
    return inv * r;
  }
  return d1c(__ldg(a + x - 2 * s), __ldg(a + x - s), __ldg(a + x + s), __ldg(a + x + 2 * s), inv);
}
__device__ __forceinline__ double gd2(const Closures &cl, const double *a, long long x, long long s, double inv2, int dir, int idx, int n) {
  if (cl.on[dir][0] && idx < cl.nr2) {
    double r = 0.0; const long long x0 = x - idx * s;
    for (int p = 0; p < cl.np2; p++) r += cl.d2[idx * cl.np2 + p] * __ldg(a + x0 + p * s);
    return inv2 * r;
  }
  if (cl.on[dir][1] && n - 1 - idx < cl.nr2) {
    const int row = n - 1 - idx; double r = 0.0; const long long x0 = x + row * s;
    for (int p = 0; p < cl.np2; p++) r += cl.d2[row * cl.np2 + p] * __ldg(a + x0 - p * s);
    return inv2 * r;
  }
  return d2c(__ldg(a + x - 2 * s), __ldg(a + x - s), __ldg(a + x), __ldg(a + x + s), __ldg(a + x + 2 * s), inv2);
}
// d/dxi_out ( d a / dxi_in ): the outer formula (central or closure by idx_out) applied to inner derivatives
__device__ __forceinline__ double gdmix(const Closures &cl, const double *a, long long x, long long sin_, double invin, int in, int idxin, int nin,
                                        long long sout, double invout, int out, int idxout, int nout) {
  if (cl.on[out][0] && idxout < cl.nr1) {
    double r = 0.0; const long long x0 = x - idxout * sout;
    for (int p = 0; p < cl.np1; p++) r += cl.d1[idxout * cl.np1 + p] * gd1(cl, a, x0 + p * sout, sin_, invin, in, idxin, nin);
    return invout * r;
  }
  if (cl.on[out][1] && nout - 1 - idxout < cl.nr1) {
    const int row = nout - 1 - idxout; double r = 0.0; const long long x0 = x + row * sout;
    for (int p = 0; p < cl.np1; p++) r -= cl.d1[row * cl.np1 + p] * gd1(cl, a, x0 - p * sout, sin_, invin, in, idxin, nin);
    return invout * r;
  }
  return d1c(gd1(cl, a, x - 2 * sout, sin_, invin, in, idxin, nin), gd1(cl, a, x - sout, sin_, invin, in, idxin, nin),
             gd1(cl, a, x + sout, sin_, invin, in, idxin, nin), gd1(cl, a, x + 2 * sout, sin_, invin, in, idxin, nin), invout);
}

// modified Ducros sensor (shock_sensors.py:12-49) on the interior
template <int ND>
__global__ void __launch_bounds__(256) k_theta(GridDev g, FieldPtrs f, PhysConst c, Closures cl, GeneralPtrs gp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z * blockDim.z + threadIdx.z;
  if (i >= g.np[0] || j >= g.np[1] || k >= g.np[2]) return;
  const long long x = g.off + i + (ND > 1 ? j * g.s[1] : 0) + (ND > 2 ? k * g.s[2] : 0);
  const int id[3] = {i, j, k};
  double du[ND][ND];
#pragma unroll
  for (int a = 0; a < ND; a++)
#pragma unroll
    for (int b = 0; b < ND; b++)
      du[a][b] = gd1(cl, f.u[a], x, g.s[b], c.inv[b], b, id[b], g.np[b]) * (gp.D[b] ? gp.D[b][x] : 1.0);
  double div = 0.0, vort = 0.0;
#pragma unroll
  for (int a = 0; a < ND; a++) div += du[a][a];
  if (ND == 2) vort = sq(du[1][0] - du[0][1]);
  if (ND == 3) vort = sq(du[2 % ND][1] - du[1][2 % ND]) + sq(du[0][2 % ND] - du[2 % ND][0]) + sq(du[1][0] - du[0][1]);
  gp.theta[x] = (0.5 - 0.5 * tanh(2.5 + 250.0 * div)) * div * div / (c.sensor_eps + div * div + vort);
}

// General viscous terms: variable viscosity, stretched (diagonal-metric) grid, one-sided closures.
// With  d_j f = D_jj delta_j f ,  d_jj f = D_jj^2 delta_jj f + D_jj SD_jjj delta_j f ,  d_ij f = D_ii D_jj delta_out(delta_in f)
// (metric.py:72-135, opensblifunctions.py:540-549) and  S_ij = d_j u_i + d_i u_j - 2/3 delta_ij div u  (app strings, e.g.
// katzer_SBLI.py:10-14, expanded by StoreSome.py:71-161):
//   momentum_i += 1/Re [ sum_j d_j mu S_ij + mu ( sum_j d_jj u_i + 1/3 sum_j d_ij u_j ) ]
//   energy     += kq [ sum_j d_j mu d_j T + mu sum_j d_jj T ] + sum_i u_i (momentum_i term) + mu/Re sum_ij S_ij d_j u_i
// viscous terms of one point from global memory (any stencil: central or one-sided closure rows): vis[a] = momentum terms
// (body force included), *en = energy term
template <int ND>
__device__ __forceinline__ void viscous_general_point(const GridDev &g, const FieldPtrs &f, const PhysConst &c, const Closures &cl,
                                                      const GeneralPtrs &gp, long long x, const int *id, double *vis, double *en) {
  const double iRe = 1.0 / c.Re;
  const double kq = iRe * (1.0 / (c.gama - 1.0)) * (1.0 / (c.Minf * c.Minf)) * (1.0 / c.Pr);
  double Dm[ND], SDm[ND], dmu[ND], dT[ND], dxi[ND + 1][ND], du[ND][ND];
#pragma unroll
  for (int d = 0; d < ND; d++) {
    Dm[d] = gp.D[d] ? gp.D[d][x] : 1.0;
    SDm[d] = gp.SD[d] ? gp.SD[d][x] : 0.0;
    dmu[d] = c.visc_law == 0 ? 0.0 : Dm[d] * gd1(cl, gp.mu, x, g.s[d], c.inv[d], d, id[d], g.np[d]);
#pragma unroll
    for (int a = 0; a < ND; a++) { dxi[a][d] = gd1(cl, f.u[a], x, g.s[d], c.inv[d], d, id[d], g.np[d]); du[a][d] = Dm[d] * dxi[a][d]; }
    dxi[ND][d] = gd1(cl, f.T, x, g.s[d], c.inv[d], d, id[d], g.np[d]);
    dT[d] = Dm[d] * dxi[ND][d];
  }
  double div = 0.0;
#pragma unroll
  for (int a = 0; a < ND; a++) div += du[a][a];
  const double mu = gp.mu[x];
  double e = 0.0;
#pragma unroll
  for (int a = 0; a < ND; a++) {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int b = 0; b < ND; b++) {
      const double Sab = du[a][b] + du[b][a] - (a == b ? (2.0 / 3.0) * div : 0.0);
      s1 += dmu[b] * Sab;
      const double lap = Dm[b] * Dm[b] * gd2(cl, f.u[a], x, g.s[b], c.inv2[b], b, id[b], g.np[b]) + Dm[b] * SDm[b] * dxi[a][b];
      if (b == a) s2 += (4.0 / 3.0) * lap;
      else {
        s2 += lap;
        const int in = a < b ? a : b, out = a < b ? b : a;
        s2 += (1.0 / 3.0) * Dm[a] * Dm[b] * gdmix(cl, f.u[b], x, g.s[in], c.inv[in], in, id[in], g.np[in], g.s[out], c.inv[out], out, id[out], g.np[out]);
      }
      e += iRe * mu * Sab * du[a][b];
    }
    vis[a] = iRe * (s1 + mu * s2);
    const double ua = __ldg(f.u[a] + x);
    e += vis[a] * ua - c.force[a] * ua;
    vis[a] -= c.force[a];
  }
  double hT = 0.0;
#pragma unroll
  for (int d = 0; d < ND; d++)
    hT += dmu[d] * dT[d] + mu * (Dm[d] * Dm[d] * gd2(cl, f.T, x, g.s[d], c.inv2[d], d, id[d], g.np[d]) + Dm[d] * SDm[d] * dxi[ND][d]);
  *en = kq * hT + e;
}

template <int ND>
__global__ void __launch_bounds__(256) k_viscous_general(GridDev g, FieldPtrs f, PhysConst c, Closures cl, GeneralPtrs gp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z * blockDim.z + threadIdx.z;
  if (i >= g.np[0] || j >= g.np[1] || k >= g.np[2]) return;
  const long long x = g.off + i + (ND > 1 ? j * g.s[1] : 0) + (ND > 2 ? k * g.s[2] : 0);
  const int id[3] = {i, j, k};
  double vis[ND], en;
  viscous_general_point<ND>(g, f, c, cl, gp, x, id, vis, &en);
  double old[ND + 1];
#pragma unroll
  for (int a = 0; a < ND + 1; a++) old[a] = f.R[1 + a][x];
  if (gp.src) f.R[0][x] += gp.src[x] * c.src_factor;      // time-periodic mass source (transitional_SBLI.py:77-89)
#pragma unroll
  for (int a = 0; a < ND; a++) f.R[1 + a][x] = old[a] + vis[a];
  f.R[ND + 1][x] = old[ND] + en;
}

// 3-D general viscous terms, tiled: a block marches along z over a 32 x 8 column with planes k-2..k+2 of (u0, u1, u2, T, mu)
// in shared memory.  Points whose stencils are all central (not within the closure rows of a flagged face) take every
// derivative from shared memory; the few rows next to walls evaluate the generic per-point formulas from global memory.
constexpr size_t vtg_smem_bytes() { return sizeof(double) * 5 * 5 * VT_PLANE; }
template <int RK>   // 0 = Residual += viscous terms, 1 / 2 = also the low-storage / SBLI RK update of the stage (in place: q is not read by stencils here)
__global__ void __launch_bounds__(VT_X * VT_Y, 2) k_viscous3d_tiled_general(GridDev g, FieldPtrs f, PhysConst c, Closures cl, GeneralPtrs gp, double rkA, double rkB) {
  extern __shared__ double vtg_smem[];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * VT_X + tx;
  const int i0 = blockIdx.x * VT_X, j0 = blockIdx.y * VT_Y, k0 = blockIdx.z * g.zlen;
  const int i = i0 + tx, j = j0 + ty;
  const int kend = min(k0 + g.zlen, g.np[2]);
  const double *src[5] = {f.u[0], f.u[1], f.u[2], f.T, gp.mu};
  auto S = [&](int slot, int v) -> double * { return vtg_smem + (size_t)(slot * 5 + v) * VT_PLANE; };
  auto load_plane = [&](int kk) {
    const int slot = (kk + 10) % 5;
    for (int e = tid; e < VT_PLANE; e += VT_X * VT_Y) {
      const int yy = e / VT_HX, xx = e % VT_HX;
      const int gi = i0 + xx - 2, gj = j0 + yy - 2;
      if (gi < g.np[0] + 2 && gj < g.np[1] + 2) {
        const long long x = g.off + gi + gj * g.s[1] + (long long)kk * g.s[2];
#pragma unroll
        for (int v = 0; v < 5; v++) S(slot, v)[e] = __ldg(src[v] + x);
      }
    }
  };
  for (int kk = k0 - 2; kk <= k0 + 2; kk++) load_plane(kk);
  const double iRe = 1.0 / c.Re;
  const double kq = iRe * (1.0 / (c.gama - 1.0)) * (1.0 / (c.Minf * c.Minf)) * (1.0 / c.Pr);
  const int ce = (ty + 2) * VT_HX + tx + 2;
  const bool active = i < g.np[0] && j < g.np[1];
  auto near_face = [&](int d, int idx) { return (cl.on[d][0] && idx < cl.nr1) || (cl.on[d][1] && g.np[d] - 1 - idx < cl.nr1); };
  const bool slow_xy = near_face(0, i) || near_face(1, j);
  __syncthreads();
  for (int k = k0; k < kend; k++) {
    const long long x = g.off + i + j * g.s[1] + (long long)k * g.s[2];
    if (active) {
      double vis[3], en;
      if (slow_xy || near_face(2, k)) {
        const int id[3] = {i, j, k};
        viscous_general_point<3>(g, f, c, cl, gp, x, id, vis, &en);
      } else {
        const double *P[5][5];
#pragma unroll
        for (int dz = 0; dz < 5; dz++)
#pragma unroll
          for (int v = 0; v < 5; v++) P[dz][v] = S((k + dz - 2 + 10) % 5, v) + ce;
        double Dm[3], SDm[3];
#pragma unroll
        for (int d = 0; d < 3; d++) { Dm[d] = gp.D[d] ? gp.D[d][x] : 1.0; SDm[d] = gp.SD[d] ? gp.SD[d][x] : 0.0; }
        double dxi[5][3], d2[4][3];
#pragma unroll
        for (int v = 0; v < 5; v++) {
          const double *p = P[2][v];
          dxi[v][0] = d1c(p[-2], p[-1], p[1], p[2], c.inv[0]);
          dxi[v][1] = d1c(p[-2 * VT_HX], p[-VT_HX], p[VT_HX], p[2 * VT_HX], c.inv[1]);
          dxi[v][2] = d1c(P[0][v][0], P[1][v][0], P[3][v][0], P[4][v][0], c.inv[2]);
          if (v < 4) {
            d2[v][0] = d2c(p[-2], p[-1], p[0], p[1], p[2], c.inv2[0]);
            d2[v][1] = d2c(p[-2 * VT_HX], p[-VT_HX], p[0], p[VT_HX], p[2 * VT_HX], c.inv2[1]);
            d2[v][2] = d2c(P[0][v][0], P[1][v][0], p[0], P[3][v][0], P[4][v][0], c.inv2[2]);
          }
        }
        // mixed xi-derivatives: outer (higher direction) central difference of the inner (lower direction) one
        auto mix = [&](int v, int in, int out) {
          double r[4];
          if (out == 1) {            // d/dy ( d/dx )
            const double *p = P[2][v];
            const int oy[4] = {-2 * VT_HX, -VT_HX, VT_HX, 2 * VT_HX};
#pragma unroll
            for (int q = 0; q < 4; q++) r[q] = d1c(p[oy[q] - 2], p[oy[q] - 1], p[oy[q] + 1], p[oy[q] + 2], c.inv[0]);
            return d1c(r[0], r[1], r[2], r[3], c.inv[1]);
          }
          const int pz[4] = {0, 1, 3, 4};
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const double *p = P[pz[q]][v];
            r[q] = in == 0 ? d1c(p[-2], p[-1], p[1], p[2], c.inv[0]) : d1c(p[-2 * VT_HX], p[-VT_HX], p[VT_HX], p[2 * VT_HX], c.inv[1]);
          }
          return d1c(r[0], r[1], r[2], r[3], c.inv[2]);
        };
        double du[3][3], dT[3], dmu[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
#pragma unroll
          for (int a = 0; a < 3; a++) du[a][d] = Dm[d] * dxi[a][d];
          dT[d] = Dm[d] * dxi[3][d];
          dmu[d] = c.visc_law == 0 ? 0.0 : Dm[d] * dxi[4][d];
        }
        const double div = du[0][0] + du[1][1] + du[2][2];
        const double mu = P[2][4][0];
        double e = 0.0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int b = 0; b < 3; b++) {
            const double Sab = du[a][b] + du[b][a] - (a == b ? (2.0 / 3.0) * div : 0.0);
            s1 += dmu[b] * Sab;
            const double lap = Dm[b] * Dm[b] * d2[a][b] + Dm[b] * SDm[b] * dxi[a][b];
            if (b == a) s2 += (4.0 / 3.0) * lap;
            else {
              s2 += lap;
              s2 += (1.0 / 3.0) * Dm[a] * Dm[b] * mix(b, a < b ? a : b, a < b ? b : a);
            }
            e += iRe * mu * Sab * du[a][b];
          }
          vis[a] = iRe * (s1 + mu * s2);
          const double ua = P[2][a][0];
          e += vis[a] * ua - c.force[a] * ua;
          vis[a] -= c.force[a];
        }
        double hT = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) hT += dmu[d] * dT[d] + mu * (Dm[d] * Dm[d] * d2[3][d] + Dm[d] * SDm[d] * dxi[3][d]);
        en = kq * hT + e;
      }
      double Rm[5];
#pragma unroll
      for (int m = 0; m < 5; m++) Rm[m] = f.R[m][x];
      if (gp.src) Rm[0] += gp.src[x] * c.src_factor;
#pragma unroll
      for (int a = 0; a < 3; a++) Rm[1 + a] += vis[a];
      Rm[4] += en;
      if (RK == 0) {
        if (gp.src) f.R[0][x] = Rm[0];
#pragma unroll
        for (int m = 1; m < 5; m++) f.R[m][x] = Rm[m];
      } else {
        double o[5], q[5];
#pragma unroll
        for (int m = 0; m < 5; m++) { o[m] = f.rk[m][x]; q[m] = f.q[m][x]; }
#pragma unroll
        for (int m = 0; m < 5; m++) {
          if (RK == 1) { const double t = c.dt * Rm[m] + rkA * o[m]; f.rk[m][x] = t; f.q[m][x] = rkB * t + q[m]; }
          else { f.q[m][x] = c.dt * rkB * Rm[m] + o[m]; f.rk[m][x] = c.dt * rkA * Rm[m] + o[m]; }
        }
      }
    }
    __syncthreads();
    if (k + 1 < kend) load_plane(k + 3);
    __syncthreads();
  }
}

// General Central(4) convective terms: one-sided closures by grid index (opensblifunctions.py:523-534), diagonal
// metrics, and the two splittings of the shipped apps: FORM 0 Blaisdell skew form (parsing.py:75-111), FORM 1 Feiereisen
// quadratic split (compressible_TCF_Central/turbulent_channel.py:12-20).  Writes Residual (first spatial kernel of a stage).
__device__ __forceinline__ int gd1_stencil(const Closures &cl, int dir, int idx, int n, double *w, int *off) {
  if (cl.on[dir][0] && idx < cl.nr1) {
    for (int p = 0; p < cl.np1; p++) { w[p] = cl.d1[idx * cl.np1 + p]; off[p] = p - idx; }
    return cl.np1;
  }
  if (cl.on[dir][1] && n - 1 - idx < cl.nr1) {
    const int row = n - 1 - idx;
    for (int p = 0; p < cl.np1; p++) { w[p] = -cl.d1[row * cl.np1 + p]; off[p] = row - p; }
    return cl.np1;
  }
  w[0] = 1.0 / 12.0; w[1] = -8.0 / 12.0; w[2] = 8.0 / 12.0; w[3] = -1.0 / 12.0;
  off[0] = -2; off[1] = -1; off[2] = 1; off[3] = 2;
  return 4;
}

template <int ND, int FORM>
__global__ void __launch_bounds__(256) k_central_general(GridDev g, FieldPtrs f, PhysConst c, Closures cl, GeneralPtrs gp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z * blockDim.z + threadIdx.z;
  if (i >= g.np[0] || j >= g.np[1] || k >= g.np[2]) return;
  const long long x = g.off + i + (ND > 1 ? j * g.s[1] : 0) + (ND > 2 ? k * g.s[2] : 0);
  const int id[3] = {i, j, k};
  constexpr int NV = ND + 2;
  double qc[NV], uc[ND], r[NV];
#pragma unroll
  for (int m = 0; m < NV; m++) { qc[m] = __ldg(f.q[m] + x); r[m] = 0.0; }
#pragma unroll
  for (int a = 0; a < ND; a++) uc[a] = __ldg(f.u[a] + x);
#pragma unroll
  for (int d = 0; d < ND; d++) {
    double w[6]; int off[6];
    const int cnt = gd1_stencil(cl, d, id[d], g.np[d], w, off);
    const double sc = c.inv[d] * (gp.D[d] ? gp.D[d][x] : 1.0);
    double dqu[NV], dq[NV], du[ND], dp = 0.0, dpu = 0.0, dh = 0.0;
#pragma unroll
    for (int m = 0; m < NV; m++) { dqu[m] = 0.0; dq[m] = 0.0; }
#pragma unroll
    for (int a = 0; a < ND; a++) du[a] = 0.0;
    for (int p = 0; p < cnt; p++) {
      const long long xs = x + off[p] * g.s[d];
      const double wp = w[p], ud = __ldg(f.u[d] + xs), pr = __ldg(f.p + xs);
      double qs[NV];
#pragma unroll
      for (int m = 0; m < NV; m++) { qs[m] = __ldg(f.q[m] + xs); dqu[m] += wp * (qs[m] * ud); dq[m] += wp * qs[m]; }
#pragma unroll
      for (int a = 0; a < ND; a++) du[a] += wp * __ldg(f.u[a] + xs);
      dp += wp * pr; dpu += wp * (pr * ud);
      if (FORM == 1) dh += wp * (qs[NV - 1] / qs[0]);
    }
#pragma unroll
    for (int m = 0; m < NV; m++) { dqu[m] *= sc; dq[m] *= sc; }
#pragma unroll
    for (int a = 0; a < ND; a++) du[a] *= sc;
    dp *= sc; dpu *= sc; dh *= sc;
    if (FORM == 0) {
#pragma unroll
      for (int m = 0; m < NV; m++) r[m] -= 0.5 * (dqu[m] + uc[d] * dq[m] + qc[m] * du[d]);
    } else {
      r[0] -= dq[1 + d];
#pragma unroll
      for (int a = 0; a < ND; a++) r[1 + a] -= 0.5 * (dqu[1 + a] + qc[1 + d] * du[a] + uc[a] * dq[1 + d]);
      r[NV - 1] -= 0.5 * (dqu[NV - 1] + qc[1 + d] * dh + (qc[NV - 1] / qc[0]) * dq[1 + d]);
    }
    r[1 + d] -= dp;
    r[NV - 1] -= dpu;
  }
#pragma unroll
  for (int m = 0; m < NV; m++) f.R[m][x] = r[m];
}

// -------------------------------------------------------------------------------------------------
// Fully curvilinear 2-D grids in strong-conservation form (apps/euler_wave_curvilinear/euler_wave.py:12-18):
//   d q / dt = - (1/detJ) sum_i  d/dxi_i [ detJ ( U_i q + p (0, D_i0, D_i1, U_i) ) ] ,   U_i = D_ij u_j
// Characteristic LLF flux with the metric-aware eigensystem (euler_eigensystem.py:18-105): direction cosines
// k~ = avg(D_i.)/|avg(D_i.)| of the two interface points, wave speeds U, U +- a |D_i.| at every stencil point
// (shock_capturing.py:357-536).  One thread per interface; the fluxes go through a work array and a difference kernel.
// -------------------------------------------------------------------------------------------------
struct CurvPtrs {
  const double *D[2][2];    // D_ij = d xi_i / d x_j
  const double *detJ;
  double *wk[4];            // interface fluxes of the direction being swept
};

template <int DIR, int RECON, int AVG>
__global__ void __launch_bounds__(128) k_flux_curv2d(GridDev g, FieldPtrs f, PhysConst c, SchemeParams sp, CurvPtrs cp) {
  // interfaces i+1/2 for i = -1 .. np[DIR]-1 along DIR, all interior points of the other direction
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  const int na = g.np[DIR] + 1;
  if (a >= na || b >= g.np[1 - DIR]) return;
  const int i = DIR == 0 ? a - 1 : b, j = DIR == 0 ? b : a - 1;
  const long long x = g.off + i + j * g.s[1], sd = g.s[DIR];
  const double gm1 = c.gama - 1.0;
  double rho_[6], m0[6], m1[6], E_[6], pr[6], u0[6], u1[6], as[6], D0[6], D1[6], J[6];
#pragma unroll
  for (int p = 0; p < 6; p++) {
    const long long xp = x + (p - 2) * sd;
    rho_[p] = __ldg(f.q[0] + xp); m0[p] = __ldg(f.q[1] + xp); m1[p] = __ldg(f.q[2] + xp); E_[p] = __ldg(f.q[3] + xp);
    D0[p] = __ldg(cp.D[DIR][0] + xp); D1[p] = __ldg(cp.D[DIR][1] + xp); J[p] = __ldg(cp.detJ + xp);
    u0[p] = m0[p] / rho_[p]; u1[p] = m1[p] / rho_[p];
    pr[p] = gm1 * (E_[p] - 0.5 * rho_[p] * (u0[p] * u0[p] + u1[p] * u1[p]));
    as[p] = sqrt(c.gama * pr[p] / rho_[p]);
  }
  // interface state (averaging.py:31-114)
  double rho, v0, v1, av;
  if (AVG == AVG_ROE) {
    const double sl = sqrt(rho_[2]), sr = sqrt(rho_[3]), w = 1.0 / (sr + sl);
    rho = sqrt(rho_[2] * rho_[3]);
    v0 = w * (sr * u0[3] + sl * u0[2]); v1 = w * (sr * u1[3] + sl * u1[2]);
    const double H = w * ((pr[2] + E_[2]) / sl + (pr[3] + E_[3]) / sr);
    av = sqrt(gm1 * (H - 0.5 * (v0 * v0 + v1 * v1)));
  } else {
    rho = 0.5 * (rho_[2] + rho_[3]); v0 = 0.5 * (u0[2] + u0[3]); v1 = 0.5 * (u1[2] + u1[3]); av = 0.5 * (as[2] + as[3]);
  }
  double k0 = 0.5 * (D0[2] + D0[3]), k1 = 0.5 * (D1[2] + D1[3]);
  const double inm = 1.0 / sqrt(k0 * k0 + k1 * k1);
  k0 *= inm; k1 *= inm;
  const double phi = 0.5 * gm1 * (v0 * v0 + v1 * v1), ia2 = 1.0 / (av * av), irho = 1.0 / rho;
  const double bt = 0.70710678118654752440 * irho / av;
  double cf[4][6], cs[4][6], lam0 = 0.0, lamp = 0.0, lamm = 0.0;
#pragma unroll
  for (int p = 0; p < 6; p++) {
    const double U = D0[p] * u0[p] + D1[p] * u1[p];
    const double am = sqrt(D0[p] * D0[p] + D1[p] * D1[p]) * as[p];
    lam0 = fmax(lam0, fabs(U)); lamp = fmax(lamp, fabs(U + am)); lamm = fmax(lamm, fabs(U - am));
    double v[2][4];
    v[0][0] = rho_[p]; v[0][1] = m0[p]; v[0][2] = m1[p]; v[0][3] = E_[p];
    v[1][0] = J[p] * (rho_[p] * U); v[1][1] = J[p] * (m0[p] * U + D0[p] * pr[p]); v[1][2] = J[p] * (m1[p] * U + D1[p] * pr[p]);
    v[1][3] = J[p] * ((pr[p] + E_[p]) * U);
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const double *xv = v[t];
      const double um = v0 * xv[1] + v1 * xv[2], w0 = xv[1] - v0 * xv[0], w1 = xv[2] - v1 * xv[0];
      const double S = phi * xv[0] - gm1 * um + gm1 * xv[3];
      const double aw = av * (k0 * w0 + k1 * w1);
      double ch[4];
      ch[0] = xv[0] - S * ia2;
      ch[1] = (k1 * w0 - k0 * w1) * irho;
      ch[2] = bt * (S + aw);
      ch[3] = bt * (S - aw);
#pragma unroll
      for (int jj = 0; jj < 4; jj++) { if (t == 0) cs[jj][p] = ch[jj]; else cf[jj][p] = ch[jj]; }
    }
  }
  double rec[4];
  rec[0] = reconstruct<RECON>(cf[0], cs[0], lam0, sp);
  rec[1] = reconstruct<RECON>(cf[1], cs[1], lam0, sp);
  rec[2] = reconstruct<RECON>(cf[2], cs[2], lamp, sp);
  rec[3] = reconstruct<RECON>(cf[3], cs[3], lamm, sp);
  const double al = 0.70710678118654752440 * rho / av;
  const double th = k0 * v0 + k1 * v1, Hp = (phi + av * av) / gm1;
  cp.wk[0][x] = rec[0] + al * (rec[2] + rec[3]);
  cp.wk[1][x] = v0 * rec[0] + k1 * rho * rec[1] + al * ((v0 + k0 * av) * rec[2] + (v0 - k0 * av) * rec[3]);
  cp.wk[2][x] = v1 * rec[0] - k0 * rho * rec[1] + al * ((v1 + k1 * av) * rec[2] + (v1 - k1 * av) * rec[3]);
  cp.wk[3][x] = (phi / gm1) * rec[0] + rho * (k1 * v0 - k0 * v1) * rec[1] + al * ((Hp + av * th) * rec[2] + (Hp - av * th) * rec[3]);
}

// Residual (+)= -(F_{i+1/2} - F_{i-1/2}) / Delta_DIR / detJ   (shock_capturing.py:21-34 with the 1/detJ of the app's equations)
template <int DIR, bool ACCUM>
__global__ void __launch_bounds__(256) k_resid_curv2d(GridDev g, FieldPtrs f, PhysConst c, CurvPtrs cp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.np[0] || j >= g.np[1]) return;
  const long long x = g.off + i + j * g.s[1];
  const double sc = -c.inv[DIR] / __ldg(cp.detJ + x);
  double r[4];
#pragma unroll
  for (int m = 0; m < 4; m++) r[m] = sc * (cp.wk[m][x] - cp.wk[m][x - g.s[DIR]]) + (ACCUM ? f.R[m][x] : 0.0);
#pragma unroll
  for (int m = 0; m < 4; m++) f.R[m][x] = r[m];
}

// -------------------------------------------------------------------------------------------------
// Boundary conditions
// -------------------------------------------------------------------------------------------------
struct Box { int lo[3], n[3]; };   // start index and extent per dimension (inactive dims: lo 0, n 1)

// periodic slab copy (periodic.py:42-56): dst box <- src box, same extents, nv arrays
__global__ void __launch_bounds__(256) k_copy_box(GridDev g, FieldPtrs f, int nv, Box src, Box dst) {
  const long long cnt = (long long)src.n[0] * src.n[1] * src.n[2];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cnt * nv; e += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(e / cnt);
    long long r = e % cnt;
    const int i = (int)(r % src.n[0]); r /= src.n[0];
    const int j = (int)(r % src.n[1]);
    const int k = (int)(r / src.n[1]);
    const long long xs = g.off + (src.lo[0] + i) + (src.lo[1] + j) * g.s[1] * (g.nd > 1) + (src.lo[2] + k) * g.s[2] * (g.nd > 2);
    const long long xd = g.off + (dst.lo[0] + i) + (dst.lo[1] + j) * g.s[1] * (g.nd > 1) + (dst.lo[2] + k) * g.s[2] * (g.nd > 2);
    f.q[m][xd] = f.q[m][xs];
  }
}
struct DirichletState { double q[5]; };
__global__ void __launch_bounds__(256) k_fill_box(GridDev g, FieldPtrs f, int nv, Box dst, DirichletState st) {
  const long long cnt = (long long)dst.n[0] * dst.n[1] * dst.n[2];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cnt * nv; e += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(e / cnt);
    long long r = e % cnt;
    const int i = (int)(r % dst.n[0]); r /= dst.n[0];
    const int j = (int)(r % dst.n[1]);
    const int k = (int)(r / dst.n[1]);
    const long long xd = g.off + (dst.lo[0] + i) + (dst.lo[1] + j) * g.s[1] * (g.nd > 1) + (dst.lo[2] + k) * g.s[2] * (g.nd > 2);
    f.q[m][xd] = st.q[m];
  }
}


// Plane boundary kernels: one thread per point of the boundary plane of (dir, side); the tangential range covers the
// scheme halos, as in the reference (bc_core.py:158-198).  `lo`/`n` = tangential start and extents (n[dir] = 1).
struct PlaneSpec { int dir, side, lo[3], n[3], nh; };   // nh = number of halo planes on that side

__device__ __forceinline__ bool plane_point(const GridDev &g, const PlaneSpec &ps, long long &x, long long &tlin) {
  const long long cnt = (long long)ps.n[0] * ps.n[1] * ps.n[2];
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cnt) return false;
  long long r = e;
  int id[3];
  id[0] = ps.lo[0] + (int)(r % ps.n[0]); r /= ps.n[0];
  id[1] = ps.lo[1] + (int)(r % ps.n[1]);
  id[2] = ps.lo[2] + (int)(r / ps.n[1]);
  x = g.off + id[0] + (g.nd > 1 ? id[1] * g.s[1] : 0) + (g.nd > 2 ? id[2] * g.s[2] : 0);
  // tangential linear index into a face table: padded index with dimension dir removed
  long long acc = 1; tlin = 0;
  for (int d = 0; d < g.nd; d++) if (d != ps.dir) { tlin += acc * (id[d] + g.h); acc *= g.pd[d]; }
  return true;
}

// dirichlet.py:28-41 with a state that varies along the face: table[m * tsize + tlin]
// free: bit m = variable m keeps its value; bit 8 = imposed energy is table + 1/2 sum(free momentum^2)/rho (transitional_SBLI.py:134-139)
__global__ void __launch_bounds__(128) k_bc_dirichlet_field(GridDev g, FieldPtrs f, int nv, PlaneSpec ps, const double *table, long long tsize, int free_mask) {
  long long x, t;
  if (!plane_point(g, ps, x, t)) return;
  const long long out = (ps.side == 0 ? -1 : 1) * g.s[ps.dir];
  for (int m = 0; m < nv; m++) {
    if (free_mask >> m & 1) continue;
    const double v = table[m * tsize + t];
    for (int h = 0; h <= ps.nh; h++) f.q[m][x + h * out] = v;
  }
  if (free_mask >> 8 & 1) {
    const double e = table[(nv - 1) * tsize + t];
    for (int h = 0; h <= ps.nh; h++) {
      const long long xo = x + h * out;
      double ke = 0.0;
      for (int m = 1; m < nv - 1; m++) if (free_mask >> m & 1) { const double v = f.q[m][xo]; ke += v * v; }
      f.q[nv - 1][xo] = e + 0.5 * ke / f.q[0][xo];
    }
  }
}
// extrapolation.py:29-58
__global__ void __launch_bounds__(128) k_bc_extrapolation(GridDev g, FieldPtrs f, int nv, PlaneSpec ps, int order) {
  long long x, t;
  if (!plane_point(g, ps, x, t)) return;
  const long long out = (ps.side == 0 ? -1 : 1) * g.s[ps.dir], in = -out;
  for (int m = 0; m < nv; m++) {
    if (order == 0) {
      const double v = f.q[m][x + in];
      for (int h = 0; h <= ps.nh; h++) f.q[m][x + h * out] = v;
    } else {
      double a = f.q[m][x - out], b = f.q[m][x];
      for (int h = 1; h <= ps.nh; h++) { const double v = 2.0 * b - a; f.q[m][x + h * out] = v; a = b; b = v; }
    }
  }
}
// inlet_pressure_extrapolate.py:32-66 (side 0)
template <int ND>
__global__ void __launch_bounds__(128) k_bc_inlet_pressure(GridDev g, FieldPtrs f, PhysConst c, PlaneSpec ps) {
  long long x, t;
  if (!plane_point(g, ps, x, t)) return;
  const long long sd = g.s[ps.dir];
  const double rhob = f.q[0][x];
  double ub[ND], ke = 0.0;
#pragma unroll
  for (int d = 0; d < ND; d++) { ub[d] = fabs(f.q[1 + d][x] / rhob); ke += ub[d] * ub[d]; }
  const double pb = (c.gama - 1.0) * (-0.5 * rhob * ke + f.q[ND + 1][x]);
  const double ab = sqrt(c.gama * pb / rhob);
  double un = ub[0];
#pragma unroll
  for (int d = 0; d < ND; d++) if (d == ps.dir) un = ub[d];
  const bool sup = un >= ab;
  if (sup) {
#pragma unroll
    for (int m = 0; m < ND + 2; m++) f.q[m][x] = f.q[m][x - sd];
  } else {
    const double E = f.q[ND + 1][x];
    for (int h = 1; h <= ps.nh; h++) f.q[ND + 1][x - h * sd] = E;
  }
}
// isothermal_wall.py:32-88
template <int ND>
__global__ void __launch_bounds__(128) k_bc_isothermal_wall(GridDev g, FieldPtrs f, PhysConst c, PlaneSpec ps) {
  long long x, t;
  if (!plane_point(g, ps, x, t)) return;
  const long long out = (ps.side == 0 ? -1 : 1) * g.s[ps.dir], in = -out;
  const double gm = c.gama, M2 = c.Minf * c.Minf;
  const double rw = f.q[0][x];
#pragma unroll
  for (int d = 0; d < ND; d++) f.q[1 + d][x] = 0.0;
  const double Ew = rw * c.Twall / (gm * (gm - 1.0) * M2);
  f.q[ND + 1][x] = Ew;
  const double Pw = (gm - 1.0) * (-0.0 / rw + Ew);
  const long long xa = x + in;
  double kea = 0.0;
  const double ra = f.q[0][xa];
#pragma unroll
  for (int d = 0; d < ND; d++) { const double m = f.q[1 + d][xa]; kea += 0.5 * m * m; }
  const double Ta = M2 * gm * (gm - 1.0) * (-kea / ra + f.q[ND + 1][xa]) / ra;
  for (int h = 1; h <= ps.nh; h++) {
    const long long xi = x + h * in, xo = x + h * out;
    const double Th = (h + 1) * c.Twall - h * Ta;
    const double rh = M2 * gm * Pw / Th;
    const double ri = f.q[0][xi];
    double u2 = 0.0, uu[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) { uu[d] = f.q[1 + d][xi] / ri; u2 += uu[d] * uu[d]; }
    f.q[0][xo] = rh;
#pragma unroll
    for (int d = 0; d < ND; d++) f.q[1 + d][xo] = -rh * uu[d];
    f.q[ND + 1][xo] = Pw / (gm - 1.0) + 0.5 * rh * u2;
  }
}
// adiabatic_wall.py:28-79: dT/dn = 0 to fourth order (T_wall from the three points above the wall), mirrored halos
template <int ND>
__global__ void __launch_bounds__(128) k_bc_adiabatic_wall(GridDev g, FieldPtrs f, PhysConst c, PlaneSpec ps) {
  long long x, t;
  if (!plane_point(g, ps, x, t)) return;
  const long long out = (ps.side == 0 ? -1 : 1) * g.s[ps.dir], in = -out;
  const double gm = c.gama, M2 = c.Minf * c.Minf;
  double T[4];
  for (int h = 1; h <= ps.nh || h <= 3; h++) {
    const long long xi = x + h * in, xo = x + h * out;
    const double r = f.q[0][xi], E = f.q[ND + 1][xi];
    double ke = 0.0, m[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) { m[d] = f.q[1 + d][xi]; ke += 0.5 * m[d] * m[d]; }
    if (h <= 3) T[h] = gm * M2 * (gm - 1.0) * (E - ke / r) / r;
    if (h <= ps.nh) {
      f.q[0][xo] = r; f.q[ND + 1][xo] = E;
#pragma unroll
      for (int d = 0; d < ND; d++) f.q[1 + d][xo] = -m[d];
    }
  }
  const double Tw = (6.0 / 11.0) * (3.0 * T[1] + (1.0 / 3.0) * T[3] - 1.5 * T[2]);
#pragma unroll
  for (int d = 0; d < ND; d++) f.q[1 + d][x] = 0.0;
  f.q[ND + 1][x] = f.q[0][x] * Tw / (gm * (gm - 1.0) * M2);
}
// symmetry.py:23-50 (cartesian normal)
__global__ void __launch_bounds__(128) k_bc_symmetry(GridDev g, FieldPtrs f, int nv, PlaneSpec ps) {
  long long x, t;
  if (!plane_point(g, ps, x, t)) return;
  const long long out = (ps.side == 0 ? -1 : 1) * g.s[ps.dir], in = -out;
  for (int h = 1; h <= ps.nh; h++)
    for (int m = 0; m < nv; m++) {
      const double v = f.q[m][x + h * in];
      f.q[m][x + h * out] = (m == 1 + ps.dir) ? v - 2.0 * v : v;
    }
}

// zero_gradient_outlet.py:12-23: boundary point <- one point inside, halos mirror the interior
__global__ void __launch_bounds__(128) k_bc_zero_gradient(GridDev g, FieldPtrs f, int nv, PlaneSpec ps) {
  long long x, t;
  if (!plane_point(g, ps, x, t)) return;
  const long long out = (ps.side == 0 ? -1 : 1) * g.s[ps.dir], in = -out;
  for (int m = 0; m < nv; m++) {
    f.q[m][x] = f.q[m][x + in];
    for (int h = 1; h <= ps.nh; h++) f.q[m][x + h * out] = f.q[m][x + h * in];
  }
}
// pressure_outlet.py:33-52: rho, momentum of the point one inside -> boundary point and halos; energy from the back pressure
template <int ND>
__global__ void __launch_bounds__(128) k_bc_pressure_outlet(GridDev g, FieldPtrs f, PhysConst c, PlaneSpec ps, double back_pressure) {
  long long x, t;
  if (!plane_point(g, ps, x, t)) return;
  const long long out = (ps.side == 0 ? -1 : 1) * g.s[ps.dir], xi = x - out;
  const double rho = f.q[0][xi];
  double m[ND], mm = 0.0;
#pragma unroll
  for (int d = 0; d < ND; d++) { m[d] = f.q[1 + d][xi]; mm += m[d] * m[d]; }
  const double E = back_pressure / (c.gama - 1.0) + 0.5 * mm / rho;
  for (int h = 0; h <= ps.nh; h++) {
    const long long xo = x + h * out;
    f.q[0][xo] = rho;
#pragma unroll
    for (int d = 0; d < ND; d++) f.q[1 + d][xo] = m[d];
    f.q[ND + 1][xo] = E;
  }
}
// inviscid_wall.py:24-52 (Cartesian normal): mirrored halos with the normal momentum reversed; boundary point <- one point
// inside with the normal momentum removed
__global__ void __launch_bounds__(128) k_bc_inviscid_wall(GridDev g, FieldPtrs f, int nv, PlaneSpec ps) {
  long long x, t;
  if (!plane_point(g, ps, x, t)) return;
  const long long out = (ps.side == 0 ? -1 : 1) * g.s[ps.dir], in = -out;
  for (int m = 0; m < nv; m++) {
    const bool normal = m == 1 + ps.dir;
    for (int h = 1; h <= ps.nh; h++) {
      const double v = f.q[m][x + h * in];
      f.q[m][x + h * out] = normal ? v - 2.0 * v : v;
    }
    const double v = f.q[m][x + in];
    f.q[m][x] = normal ? v - 1.0 * v : v;
  }
}

// -------------------------------------------------------------------------------------------------
// In-loop diagnostics: NaN check of a dataset (what ops_NaNcheck does, simulation_monitors.py:112-113,
// helperfunctions.py:172-190) and volume sums for the Taylor-Green diagnostics (kinetic energy, enstrophy -- computed offline
// from the dumps in the reference workflow).  Deterministic: one partial per block (warp shuffles, then one value per warp
// through shared memory), summed in block order by a second single-block kernel.
// -------------------------------------------------------------------------------------------------
constexpr int DIAG_N = 6;     // sum rho, sum 1/2 rho |u|^2, sum 1/2 rho |omega|^2, sum rhoE, max |u|/a (Mach), non-finite count
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = dmax2(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int ND>
__global__ void __launch_bounds__(256) k_diag_partial(GridDev g, FieldPtrs f, PhysConst c, GeneralPtrs gp, const double *field, double *partial) {
  __shared__ double sh[DIAG_N][8];
  const long long cnt = (long long)g.np[0] * g.np[1] * g.np[2];
  double acc[DIAG_N] = {0, 0, 0, 0, 0, 0};
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % g.np[0]);
    const long long r = e / g.np[0];
    const int j = (int)(r % g.np[1]), k = (int)(r / g.np[1]);
    const long long x = g.off + i + (ND > 1 ? j * g.s[1] : 0) + (ND > 2 ? k * g.s[2] : 0);
    if (field) {                     // NaN / Inf count of one dataset
      const double v = field[x];
      if (!(fabs(v) <= 1.79769313486231570e308)) acc[5] += 1.0;
      continue;
    }
    const double rho = f.q[0][x], irho = 1.0 / rho;
    double u2 = 0.0;
#pragma unroll
    for (int d = 0; d < ND; d++) { const double u = f.q[1 + d][x] * irho; u2 += u * u; }
    const double E = f.q[ND + 1][x];
    const double p = (c.gama - 1.0) * (E - 0.5 * rho * u2);
    // vorticity from 4th-order central differences of u = m / rho (halos valid after the boundary conditions)
    double du[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int a = 0; a < ND; a++)
#pragma unroll
      for (int b = 0; b < ND; b++) {
        if (a == b) continue;
        const long long s = g.s[b];
        const double um2 = f.q[1 + a][x - 2 * s] / f.q[0][x - 2 * s], um1 = f.q[1 + a][x - s] / f.q[0][x - s];
        const double up1 = f.q[1 + a][x + s] / f.q[0][x + s], up2 = f.q[1 + a][x + 2 * s] / f.q[0][x + 2 * s];
        du[a][b] = d1c(um2, um1, up1, up2, c.inv[b]) * (gp.D[b] ? gp.D[b][x] : 1.0);
      }
    double w2 = 0.0;
    if (ND == 2) w2 = sq(du[1][0] - du[0][1]);
    if (ND == 3) w2 = sq(du[2][1] - du[1][2]) + sq(du[0][2] - du[2][0]) + sq(du[1][0] - du[0][1]);
    acc[0] += rho; acc[1] += 0.5 * rho * u2; acc[2] += 0.5 * rho * w2; acc[3] += E;
    acc[4] = dmax2(acc[4], sqrt(u2 * rho / (c.gama * p)));
    if (!(fabs(rho) <= 1.79769313486231570e308) || !(fabs(E) <= 1.79769313486231570e308)) acc[5] += 1.0;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int n = 0; n < DIAG_N; n++) {
    const double v = n == 4 ? warp_max(acc[n]) : warp_sum(acc[n]);
    if (lane == 0) sh[n][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < DIAG_N) {
    double v = sh[threadIdx.x][0];
    for (int q = 1; q < 8; q++) v = threadIdx.x == 4 ? dmax2(v, sh[threadIdx.x][q]) : v + sh[threadIdx.x][q];
    partial[(long long)blockIdx.x * DIAG_N + threadIdx.x] = v;
  }
}
__global__ void k_diag_final(const double *partial, int nblocks, double *out) {
  const int n = threadIdx.x;
  if (n >= DIAG_N) return;
  double v = partial[n];
  for (int b = 1; b < nblocks; b++) v = n == 4 ? dmax2(v, partial[(long long)b * DIAG_N + n]) : v + partial[(long long)b * DIAG_N + n];
  out[n] = v;
}

// FP64 pipe micro-benchmark: 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace osb
