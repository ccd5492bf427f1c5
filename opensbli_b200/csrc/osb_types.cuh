// osb_types.cuh -- argument structures shared by every kernel translation unit, and the staging helpers of the flux sweeps.
#pragma once
#include "osb_math.cuh"
#include "osb_flux.cuh"

namespace osb {

struct GridDev {
  int nd;
  int np[3];          // interior points
  int pd[3];          // padded dims (np + 2h in active dims)
  long long s[3];     // strides
  long long off;      // linear index of point (0,0,0)
  int h;              // storage halo
  int zlen;           // planes per block of the z-marching kernels (32 on large grids, shorter when that leaves SMs idle)
  long long n;        // padded size
};

// n / d for n < 2^31 as one multiply-high, one add and one shift (Granlund-Montgomery round-up method; the magic number is
// computed on the host).  The sweeps number their points / rows consecutively across pencils and would otherwise spend ~100
// instructions per 64-bit division on the way back to (i, j, k).
struct FastDiv {
  unsigned m, l, d;
  FastDiv() : m(1), l(0), d(1) {}
  explicit FastDiv(unsigned div) : d(div) {
    l = 0;
    while ((1ull << l) < div) l++;
    m = (unsigned)(((1ull << 32) * ((1ull << l) - div)) / div + 1);
  }
#if defined(__CUDACC__)
  __device__ __forceinline__ unsigned div(unsigned n) const { return (__umulhi(m, n) + n) >> l; }
#endif
};

struct FieldPtrs {
  double *q[5];       // rho, rhou0.., rhoE
  double *u[3];       // velocities
  double *p, *a, *T;
  double *R[5];       // Residual
  double *rk[5];      // RK register (tempRK_* or *_RKold)
};

struct PhysConst {
  double gama, Minf, Re, Pr, dt;
  double inv[3], inv2[3];   // 1/Delta_d, 1/Delta_d^2
  // general path
  int visc_law;             // 0 constant, 1 Sutherland, 2 power law
  double SuthT, RefT, mu_exp, Twall, sensor_eps;
  double force[3];          // constant body force c_j (channel apps): momentum_i -= c_i, energy -= c_j u_j
  double src_factor;        // sin(src_rate * iteration) of the mass source, refreshed by the host every step
};

// one-sided closure tables (reduced_access_scheme.py:36-83, Carpenter_scheme.py:38-102): rows idx = 0..nr-1 next to
// side 0 x weights of the boundary-absolute points 0..np-1; side 1 mirrors them (sign -1 for first derivatives)
struct Closures {
  int on[3][2];
  int nr1, np1, nr2, np2;
  double d1[4 * 6], d2[2 * 6];
};

struct GeneralPtrs {
  const double *D[3];       // D_dd metric (nullptr: direction not stretched)
  const double *SD[3];      // SD_ddd
  double *mu, *theta, *teno_store;
  const double *src;        // mass-source amplitude (nullptr: none)
};

// constituent relations of one staged point (velocity, pressure, speed of sound of the canonical system; the app strings
// e.g. Sod_shock_tube.py:26-28) in the layout SV<ND>; DIR = sweep direction
template <int ND, int DIR>
__device__ __forceinline__ void stage_values(const double *q, double gama, double *sv, int VS) {
  typedef SV<ND> V;
  const double rho = q[0], E = q[ND + 1];
  const double y = rsqrt_nr(rho), irho = y * y;
  double mu = 0.0;                                    // m . u = 2 x kinetic energy
#pragma unroll
  for (int d = 0; d < ND; d++) {
    const double m = q[1 + d];
    sv[(V::M0 + d) * VS] = m;
    const double u = m * irho;
    if (d == DIR) sv[V::UD * VS] = u;
    mu = fma(m, u, mu);
  }
  const double p = (gama - 1.0) * (E - 0.5 * mu);
  sv[V::RHO * VS] = rho; sv[V::Y * VS] = y; sv[V::E * VS] = E; sv[V::P * VS] = p;
  const double a2 = gama * p * irho;
  sv[V::A * VS] = a2 * rsqrt_nr(a2);
}

// Loads whose values are needed only at the end of a block / plane iteration (old Residual, RK register) would expose their
// DRAM latency to every warp at once where they are used: the lines are requested into L2 up front (a few lanes per warp).
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int ND, int DIR>
__device__ __forceinline__ void stage_point(const FieldPtrs &f, long long x, double gama, double *sv, int VS) {
  double q[ND + 2];
#pragma unroll
  for (int m = 0; m < ND + 2; m++) q[m] = __ldg(f.q[m] + x);
  stage_values<ND, DIR>(q, gama, sv, VS);
}

struct alignas(64) TmaMaps5 { unsigned char m[5][128]; };     // five CUtensorMap objects (rho, rhou0, rhou1, rhou2, rhoE)

// ---- TMA + mbarrier primitives (PTX; SASS: UTMALDG / SYNCS) shared by the marching flux sweeps and the stage kernels
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void *dst, const void *map, unsigned long long *bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

}  // namespace osb
