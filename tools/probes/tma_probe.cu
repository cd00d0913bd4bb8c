// Stand-alone probe of the TMA path used by gemm.cu: one 16 (k) x 64 (rows) FP64 box with the 128-byte swizzle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, double *out, int c0, int c1, int variant) {
  __shared__ __align__(1024) unsigned char tile[8192];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
    if (variant & 1) asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    else asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(8192) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(smem_u32(tile)), "l"(&map), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
               ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const int r = i / 16, k = i % 16;
    out[i] = *reinterpret_cast<const double *>(tile + r * 128 + ((((k >> 1) ^ (r & 7))) << 4) + ((k & 1) << 3));
  }
}
int main(int argc, char **argv) {
  const int C0 = argc > 1 ? atoi(argv[1]) : 32, C1 = argc > 2 ? atoi(argv[2]) : 70;
  const int L = 256, ld = 256;
  std::vector<double> h((size_t)L * ld);
  for (int i = 0; i < L; ++i) for (int j = 0; j < ld; ++j) h[(size_t)i * ld + j] = i * 1000 + j;
  double *d, *o;
  cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, 1024 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)L, (cuuint64_t)L}; const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  const cuuint32_t box[2] = {16, 64}; const cuuint32_t es[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  for (int variant = 0; variant < 2; ++variant) {
    probe<<<1, 128>>>(map, o, C0, C1, variant);
    e = cudaDeviceSynchronize();
    printf("variant %d: %s\n", variant, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<double> got(1024);
    cudaMemcpy(got.data(), o, 1024 * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 1024; ++i) { int rr = i / 16, k = i % 16; if (got[i] != (C1 + rr) * 1000 + C0 + k) ++bad; }
    printf("variant %d mismatches %d (first %g expect %g)\n", variant, bad, got[0], C1 * 1000.0 + C0);
  }
  return 0;
}
