// osb_flux_api.h -- host-side entry points of the flux-sweep translation units (osb_flux_tu.cu, compiled once per
// dimensionality and reconstruction so that the build parallelises and a kernel experiment recompiles one small file).
#pragma once
#include <cuda_runtime.h>
#include "osb_types.cuh"

namespace osb {

struct FluxArgs {
  GridDev g;
  FieldPtrs f;
  PhysConst c;
  SchemeParams sp;
  AdaptiveCT ad;
  GeneralPtrs gp;
};

// consecutive numbering of the staged points (x) / rows (y, z) of a sweep: length of a pencil incl. 3 + 3 halo points and
// the extent of the next-faster index
struct SweepIdx {
  FastDiv len, n1;
  unsigned total;
};

// One characteristic flux sweep along `dir` (0 = x) + flux difference into Residual (accum: Residual += ...).
// Returns cudaGetLastError() of the launch.  One definition per (ndim, recon): flux3_sweep_<ndim>_<recon>.
#define OSB_FLUX_DECL(ND, R) cudaError_t flux3_sweep_##ND##_##R(int dir, int avg, bool accum, const FluxArgs &a, cudaStream_t s);
OSB_FLUX_DECL(1, 0) OSB_FLUX_DECL(1, 1) OSB_FLUX_DECL(1, 2) OSB_FLUX_DECL(1, 3)
OSB_FLUX_DECL(2, 0) OSB_FLUX_DECL(2, 1) OSB_FLUX_DECL(2, 2) OSB_FLUX_DECL(2, 3)
OSB_FLUX_DECL(3, 0) OSB_FLUX_DECL(3, 1) OSB_FLUX_DECL(3, 2) OSB_FLUX_DECL(3, 3)
#undef OSB_FLUX_DECL

inline cudaError_t flux3_sweep(int nd, int recon, int dir, int avg, bool accum, const FluxArgs &a, cudaStream_t s) {
  typedef cudaError_t (*fn_t)(int, int, bool, const FluxArgs &, cudaStream_t);
  static const fn_t table[3][4] = {{flux3_sweep_1_0, flux3_sweep_1_1, flux3_sweep_1_2, flux3_sweep_1_3},
                                   {flux3_sweep_2_0, flux3_sweep_2_1, flux3_sweep_2_2, flux3_sweep_2_3},
                                   {flux3_sweep_3_0, flux3_sweep_3_1, flux3_sweep_3_2, flux3_sweep_3_3}};
  return table[nd - 1][recon](dir, avg, accum, a, s);
}

}  // namespace osb
