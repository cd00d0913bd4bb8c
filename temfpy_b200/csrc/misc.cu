// Library identity, device queries and the FP64 peak probe used as roofline denominator.
#include "cta.hpp"

#if defined(TMF_HOSTSIM)
namespace tmfsim {
thread_local int block_id = 0;
thread_local int n_threads = 1;
thread_local unsigned char *smem = nullptr;
}  // namespace tmfsim
#endif

extern "C" int tmf_version(void) { return 100; }

extern "C" int tmf_is_cuda(void) {
#if defined(TMF_HOSTSIM)
  return 0;
#else
  return 1;
#endif
}

extern "C" int tmf_device_count(void) {
#if defined(TMF_HOSTSIM)
  return 0;
#else
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
#endif
}

#if !defined(TMF_HOSTSIM)
// 8 independent DFMA chains per thread, 256 threads per CTA, enough CTAs to fill every SM.
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double *sink) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) sink[0] = s;  // never true; keeps the chains alive
}
#endif

extern "C" int tmf_fp64_peak_probe(int iters, double *sink_dev, float *ms_out, double *flops_out,
                                   void *stream) {
#if defined(TMF_HOSTSIM)
  (void)iters; (void)sink_dev; (void)stream;
  *ms_out = 0.f;
  *flops_out = 0.0;
  tmf::set_error("fp64 probe needs a CUDA device");
  return TMF_ERR_RUNTIME;
#else
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaStream_t st = (cudaStream_t)stream;
  fp64_probe_kernel<<<grid, 256, 0, st>>>(iters / 8 + 1, sink_dev);  // warm-up
  cudaEventRecord(e0, st);
  fp64_probe_kernel<<<grid, 256, 0, st>>>(iters, sink_dev);
  cudaEventRecord(e1, st);
  cudaError_t err = cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms;
  *flops_out = 2.0 * 8.0 * (double)iters * 256.0 * (double)grid;
  return tmf::check_cuda(err, "fp64 probe");
#endif
}
