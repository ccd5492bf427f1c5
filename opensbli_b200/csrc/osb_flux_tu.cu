// osb_flux_tu.cu -- instantiations of the flux-sweep kernels for ONE (ndim, reconstruction) pair:
//   nvcc -DOSB_FLUX_ND=3 -DOSB_FLUX_RECON=2 -c osb_flux_tu.cu     (RECON: 0 WENO5-JS, 1 WENO5-Z, 2 TENO5, 3 TENO6)
#include <cuda.h>
#include <array>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include "osb_flux_api.h"
#include "osb_flux3.cuh"
#include "osb_tma_host.h"

#ifndef OSB_FLUX_ND
#error "compile with -DOSB_FLUX_ND=1|2|3 -DOSB_FLUX_RECON=0|1|2|3"
#endif

namespace osb {
namespace {

constexpr int ND = OSB_FLUX_ND, RECON = OSB_FLUX_RECON;

template <int AVG, bool ACCUM>
cudaError_t sweep_x(const FluxArgs &a, cudaStream_t s) {
  const GridDev &g = a.g;
  const long long T = (long long)(g.np[0] + 6) * g.np[1] * g.np[2];
  if (T >= (1LL << 31)) return cudaErrorInvalidValue;        // 32-bit point numbering (a 2^31-point block would not fit in HBM anyway)
  const long long nb = (T + F3_BT - 7) / (F3_BT - 6);
  SweepIdx si;
  si.len = FastDiv((unsigned)(g.np[0] + 6)); si.n1 = FastDiv((unsigned)g.np[1]); si.total = (unsigned)T;
  auto kern = k_flux3_x<ND, RECON, AVG, ACCUM>;
  constexpr size_t smem = f3_x_smem_bytes<ND, RECON>();
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)nb, F3_BT, smem, s>>>(g, a.f, a.c, a.sp, a.ad, a.gp, si);
  return cudaGetLastError();
}

#if OSB_FLUX_ND == 3
// box of 34 x-lanes (see F3_MBOX) x F3_MTY rows along DIR of one padded array (x fastest); false if TMA cannot address it
template <int DIR>
bool make_map(const GridDev &g, double *base, unsigned char *out) {
  return tma_make_map(g, base, F3_MBOX, DIR == 1 ? F3_MTY : 1, DIR == 2 ? F3_MTY : 1, out);
}

bool march_enabled() {
  static const bool on = getenv("OSB_NO_FLUX_MARCH") == nullptr;
  return on;
}

// marching TMA kernel when the layout allows it and there is enough work to fill the GPU; returns cudaErrorNotSupported otherwise
template <int DIR, int AVG, bool ACCUM>
cudaError_t sweep_march(const FluxArgs &a, cudaStream_t s) {
  const GridDev &g = a.g;
  constexpr int OTH = (DIR == 1) ? 2 : 1;
  if (!march_enabled() || a.ad.on || a.gp.D[DIR] || (g.s[1] & 1) || (g.s[2] & 1) || ((g.h - 1) & 1)) return cudaErrorNotSupported;
  const int nx = (g.np[0] + 31) / 32;
  // rows per chunk: whole pencils when they alone fill the GPU (2 blocks x 148 SMs x a few waves), else pieces of >= 64 rows
  int chunk = g.np[DIR];
  while ((long long)nx * g.np[OTH] * ((g.np[DIR] + chunk - 1) / chunk) < 148 * 2 * 4 && chunk > 64) chunk = (chunk + 1) / 2;
  if ((long long)nx * g.np[OTH] * ((g.np[DIR] + chunk - 1) / chunk) < 148 * 2) return cudaErrorNotSupported;
  TmaMaps5 maps;
  for (int m = 0; m < 5; m++) if (!make_map<DIR>(g, a.f.q[m], maps.m[m])) return cudaErrorNotSupported;
  auto kern = k_flux3_march<DIR, RECON, AVG, ACCUM>;
  constexpr size_t smem = f3_march_smem_bytes<RECON>();
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 b(32, F3_MTY, 1), gr(nx, g.np[OTH], (g.np[DIR] + chunk - 1) / chunk);
  if (gr.y > 65535u || gr.z > 65535u) return cudaErrorNotSupported;
  kern<<<gr, b, smem, s>>>(g, a.f, a.c, a.sp, maps, chunk);
  return cudaGetLastError();
}
#endif

template <int DIR, int AVG, bool ACCUM>
cudaError_t sweep_yz(const FluxArgs &a, cudaStream_t s) {
  const GridDev &g = a.g;
#if OSB_FLUX_ND == 3
  if (RECON != RECON_TENO6) {          // (TENO6 columns do not leave room for two marching blocks per SM)
    const cudaError_t e = sweep_march<DIR, AVG, ACCUM>(a, s);
    if (e != cudaErrorNotSupported) return e;
  }
#endif
  constexpr int TY = f3_ty<RECON>();
  constexpr int OTH = (DIR == 1) ? 2 : 1;
  const long long TR = (long long)(g.np[DIR] + 6) * (ND > 2 ? g.np[OTH] : 1);
  if (TR >= (1LL << 31)) return cudaErrorInvalidValue;
  SweepIdx si;
  si.len = FastDiv((unsigned)(g.np[DIR] + 6)); si.total = (unsigned)TR;
  dim3 b(32, TY, 1), gr((unsigned)((TR + TY - 2) / (TY - 1)), (g.np[0] + 31) / 32, 1);
  auto kern = k_flux3_yz<(ND >= 2 ? ND : 2), DIR, RECON, AVG, ACCUM>;
  constexpr size_t smem = f3_yz_smem_bytes<(ND >= 2 ? ND : 2), RECON>();
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<gr, b, smem, s>>>(g, a.f, a.c, a.sp, a.ad, a.gp, si);
  return cudaGetLastError();
}

}  // namespace

#define OSB_CAT3(a, b, c) a##b##_##c
#define OSB_NAME(nd, r) OSB_CAT3(flux3_sweep_, nd, r)

// Sweep order of the driver: 3-D z (first: overwrites Residual), x, y (accumulate); 2-D x, y; 1-D x.
cudaError_t OSB_NAME(OSB_FLUX_ND, OSB_FLUX_RECON)(int dir, int avg, bool accum, const FluxArgs &a, cudaStream_t s) {
  const bool roe = avg == AVG_ROE;
  if (dir == 0) {
    if (accum) return roe ? sweep_x<AVG_ROE, true>(a, s) : sweep_x<AVG_SIMPLE, true>(a, s);
    return roe ? sweep_x<AVG_ROE, false>(a, s) : sweep_x<AVG_SIMPLE, false>(a, s);
  }
#if OSB_FLUX_ND >= 2
  if (dir == 1) {
    if (accum) return roe ? sweep_yz<1, AVG_ROE, true>(a, s) : sweep_yz<1, AVG_SIMPLE, true>(a, s);
    return roe ? sweep_yz<1, AVG_ROE, false>(a, s) : sweep_yz<1, AVG_SIMPLE, false>(a, s);
  }
#endif
#if OSB_FLUX_ND >= 3
  if (dir == 2) {
    if (accum) return roe ? sweep_yz<2, AVG_ROE, true>(a, s) : sweep_yz<2, AVG_SIMPLE, true>(a, s);
    return roe ? sweep_yz<2, AVG_ROE, false>(a, s) : sweep_yz<2, AVG_SIMPLE, false>(a, s);
  }
#endif
  return cudaErrorInvalidValue;
}

}  // namespace osb
