// osb_tma_host.h -- host side of the TMA paths: tensor-map descriptors of the padded arrays (driver API entry point fetched
// through the runtime: the library does not link libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <array>
#include <cstring>
#include <map>
#include <tuple>
#include "osb_types.cuh"

namespace osb {

typedef CUresult (*encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline encode_tiled_t encode_tiled() {
  static encode_tiled_t fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (encode_tiled_t)p;
  }();
  return fn;
}

// 128-byte descriptor of a box bx x by x bz (x fastest) of one padded fp64 array; cached per (array, box).  false if TMA cannot
// address the array: global strides must be multiples of 16 bytes (even padded x-extent); the START of every box along x must be
// 16-byte aligned as well (probed on the B200: an odd start coordinate raises "illegal instruction"), which is the caller's business.
inline bool tma_make_map(const GridDev &g, double *base, int bx, int by, int bz, unsigned char *out) {
  static std::map<std::tuple<const void *, int, int, int, int, int, int>, std::array<unsigned char, 128>> cache;
  if ((g.s[1] & 1) || (g.s[2] & 1) || (bx & 1)) return false;
  const auto key = std::make_tuple((const void *)base, bx, by, bz, g.pd[0], g.pd[1], g.pd[2]);
  auto it = cache.find(key);
  if (it == cache.end()) {
    encode_tiled_t enc = encode_tiled();
    if (!enc) return false;
    alignas(64) CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)g.pd[0], (cuuint64_t)g.pd[1], (cuuint64_t)g.pd[2]};
    const cuuint64_t strides[2] = {(cuuint64_t)g.s[1] * sizeof(double), (cuuint64_t)g.s[2] * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
    std::array<unsigned char, 128> raw;
    memcpy(raw.data(), &m, 128);
    it = cache.emplace(key, raw).first;
  }
  memcpy(out, it->second.data(), 128);
  return true;
}

}  // namespace osb
