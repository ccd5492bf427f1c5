// tma_probe2.cu -- which tensor-map parameters does a 3-D/2-D TMA box load of 8-byte elements accept? (experiment harness)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap map, double *out, int c0, int c1, int c2, int bytes) {
  extern __shared__ __align__(128) double sm[];
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + 8 * 32);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(sm)), "l"(&map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(sm)), "l"(&map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
  }
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
  } while (!ok);
  for (int e = threadIdx.x; e < 8 * 32; e += blockDim.x) out[e] = sm[e];
}
int main(int argc, char **argv) {
  // args: rank dtype(0 f64, 1 u64, 2 f32 pairs) c0 l2promo oob pd0
  const int rank = atoi(argv[1]), dt = atoi(argv[2]), c0 = atoi(argv[3]), l2 = atoi(argv[4]), oob = atoi(argv[5]), pd0 = atoi(argv[6]);
  const int pd1 = 40, pd2 = 30, rows = 8;
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
  const int f = dt == 2 ? 2 : 1;
  const cuuint64_t dims[3] = {(cuuint64_t)pd0 * f, (cuuint64_t)pd1, (cuuint64_t)pd2};
  const cuuint64_t strides[2] = {(cuuint64_t)pd0 * 8, (cuuint64_t)pd0 * pd1 * 8};
  const cuuint32_t box[3] = {32u * f, (cuuint32_t)rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUtensorMapDataType dts[3] = {CU_TENSOR_MAP_DATA_TYPE_FLOAT64, CU_TENSOR_MAP_DATA_TYPE_UINT64, CU_TENSOR_MAP_DATA_TYPE_FLOAT32};
  CUresult r = enc(&m, dts[dt], rank, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   (CUtensorMapL2promotion)l2, (CUtensorMapFloatOOBfill)oob);
  printf("rank %d dtype %d c0 %d l2 %d oob %d pd0 %d: encode rc %d; ", rank, dt, c0, l2, oob, pd0, (int)r);
  const int c1 = 3, c2 = 7;
  const size_t smem = 8 * 32 * 8 + 128;
  if (rank == 3) k<3><<<1, 256, smem>>>(m, o, c0 * f, c1, c2, rows * 32 * 8); else k<2><<<1, 256, smem>>>(m, o, c0 * f, c1, 0, rows * 32 * 8);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<double> got(rows * 32);
    cudaMemcpy(got.data(), o, rows * 32 * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < rows; rr++) for (int x = 0; x < 32; x++) {
      const size_t idx = (size_t)(c0 + x) + (size_t)(c1 + rr) * pd0 + (size_t)(rank == 3 ? c2 : 0) * pd0 * pd1;
      if (got[rr * 32 + x] != (double)idx) bad++;
    }
    printf("; mismatches %d", bad);
  }
  printf("\n");
  return 0;
}
