// tma_probe.cu -- minimal TMA 3-D box load of doubles (experiment harness; not part of the product library)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu ; ./tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct alignas(64) Maps { unsigned char m[5][128]; };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__constant__ Maps cmaps;

template <int VAR>
__global__ void k(const __grid_constant__ Maps maps, double *out, int c0, int c1, int c2, int rows, const void *gmap, const double *gsrc) {
  extern __shared__ __align__(128) double sm[];
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + 5 * 8 * 32);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(rows * 32 * 8) : "memory");
    if (VAR == 16)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sm)), "l"(gsrc), "r"(rows * 32 * 8), "r"(smem_u32(bar)) : "memory");
    else if (VAR == 8) {
      asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(gmap) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(sm)), "l"(gmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    } else if (VAR == 32)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(sm)), "l"(&cmaps.m[0][0]), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else if (VAR == 0)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(sm)), "l"(maps.m[0]), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(sm)), "l"(&maps.m[0][0]), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
  } while (!ok);
  for (int e = threadIdx.x; e < rows * 32; e += blockDim.x) out[e] = sm[e];
}

int main(int argc, char **argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int pd0 = 138, pd1 = 40, pd2 = 30, rows = 8;
  const size_t n = (size_t)pd0 * pd1 * pd2;
  std::vector<double> h(n);
  for (size_t i = 0; i < n; i++) h[i] = (double)i;
  double *d, *o;
  cudaMalloc(&d, n * 8); cudaMalloc(&o, rows * 32 * 8);
  cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
  typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  enc_t enc = (enc_t)p;
  alignas(64) CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)pd0, (cuuint64_t)pd1, (cuuint64_t)pd2};
  const cuuint64_t strides[2] = {(cuuint64_t)pd0 * 8, (cuuint64_t)pd0 * pd1 * 8};
  const bool zdir = (variant & 2) != 0;
  const cuuint32_t box[3] = {32u, zdir ? 1u : (cuuint32_t)rows, zdir ? (cuuint32_t)rows : 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(&m, (variant & 4) ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode rc %d\n", variant, (int)r);
  Maps maps; memset(&maps, 0, sizeof(maps)); memcpy(maps.m[0], &m, 128);
  const int c0 = 5, c1 = zdir ? 7 : 3, c2 = zdir ? 3 : 7;
  const size_t smem = 5 * 8 * 32 * 8 + 128;
  void *gmap; cudaMalloc(&gmap, 128); cudaMemcpy(gmap, &m, 128, cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(cmaps, &maps, sizeof(maps));
  if (variant & 16) k<16><<<1, 256, smem>>>(maps, o, c0, c1, c2, rows, gmap, d + 16);
  else if (variant & 8) k<8><<<1, 256, smem>>>(maps, o, c0, c1, c2, rows, gmap, d);
  else if (variant & 32) k<32><<<1, 256, smem>>>(maps, o, c0, c1, c2, rows, gmap, d);
  else if (variant & 1) k<1><<<1, 256, smem>>>(maps, o, c0, c1, c2, rows, gmap, d); else k<0><<<1, 256, smem>>>(maps, o, c0, c1, c2, rows, gmap, d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d kernel: %s\n", variant, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<double> got(rows * 32);
    cudaMemcpy(got.data(), o, rows * 32 * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < rows; rr++) for (int x = 0; x < 32; x++) {
      const size_t idx = zdir ? (size_t)(c0 + x) + (size_t)c1 * pd0 + (size_t)(c2 + rr) * pd0 * pd1 : (size_t)(c0 + x) + (size_t)(c1 + rr) * pd0 + (size_t)c2 * pd0 * pd1;
      const double want = (variant & 16) ? (double)(16 + rr * 32 + x) : (double)idx;
      if (got[rr * 32 + x] != want) bad++;
    }
    printf("variant %d mismatches %d\n", variant, bad);
  }
  return 0;
}
