// CTA-level programming helpers shared by every kernel.
//
// Kernels are written as bulk-synchronous CTA programs: phases made of parallel-for loops over
// work items (PAR_FOR) or threads (THREAD_FOR), separated by CTA_SYNC().  With nvcc this maps to
// blockDim-strided loops and __syncthreads().  With -DTMF_HOSTSIM the same source is compiled by
// g++ into a sequential CTA simulator (one "thread" at a time, phases in program order) that the
// CPU test-suite uses to check kernel *logic* without a GPU (tests/hostsim/; never loaded by the
// package itself, which requires the CUDA build).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/temfpy_b200.h"

namespace tmf {
void set_error(const std::string &msg);
}

#if defined(TMF_HOSTSIM)
// ------------------------------------------------------------------------------------------
#define TMF_GLOBAL static void
#define TMF_GLOBAL_LB(threads, blocks) static void
#define TMF_DEVICE static inline
#define TMF_HD static inline
#define TMF_RESTRICT
namespace tmfsim {
extern thread_local int block_id;
extern thread_local int n_threads;
extern thread_local unsigned char *smem;
}  // namespace tmfsim
#define BLOCK_ID (tmfsim::block_id)
#define NTHREADS (tmfsim::n_threads)
#define PAR_FOR(i, n) for (int i = 0; i < (int)(n); ++i)
#define THREAD_FOR(t) for (int t = 0; t < tmfsim::n_threads; ++t)
#define CTA_SYNC() ((void)0)
#define DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(tmfsim::smem)

namespace tmf {
void count_launch();
template <class K, class... Args>
inline int launch_t(const char * /*tag*/, K kernel, int grid, int block, size_t smem_bytes, void * /*stream*/,
                    Args... args) {
  count_launch();
  std::vector<unsigned char> smem(smem_bytes + 64);
  tmfsim::n_threads = block;
  tmfsim::smem = smem.data();
  for (int b = 0; b < grid; ++b) {
    tmfsim::block_id = b;
    kernel(args...);
  }
  return TMF_OK;
}
template <class K, class... Args>
inline int launch(K kernel, int grid, int block, size_t smem_bytes, void *stream, Args... args) {
  return launch_t("other", kernel, grid, block, smem_bytes, stream, args...);
}
inline int copy_h2d(void *dst, const void *src, size_t bytes, void *) {
  std::memcpy(dst, src, bytes);
  return TMF_OK;
}
inline int copy_d2h_sync(void *dst, const void *src, size_t bytes, void *) {
  std::memcpy(dst, src, bytes);
  return TMF_OK;
}
inline int copy_d2h_async(void *dst, const void *src, size_t bytes, void *) {
  std::memcpy(dst, src, bytes);
  return TMF_OK;
}
inline int memset_dev(void *dst, int v, size_t bytes, void *) {
  std::memset(dst, v, bytes);
  return TMF_OK;
}
inline int stream_sync(void *) { return TMF_OK; }
}  // namespace tmf

#else
// ------------------------------------------------------------------------------------------
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#define TMF_GLOBAL __global__ void
#define TMF_GLOBAL_LB(threads, blocks) __global__ void __launch_bounds__(threads, blocks)
#define TMF_DEVICE __device__ __forceinline__
#define TMF_HD __host__ __device__ __forceinline__
#define TMF_RESTRICT __restrict__
#define BLOCK_ID ((int)blockIdx.x)
#define NTHREADS ((int)blockDim.x)
#define PAR_FOR(i, n) for (int i = threadIdx.x; i < (int)(n); i += blockDim.x)
#define THREAD_FOR(t) for (int t = threadIdx.x, t##_once = 1; t##_once; t##_once = 0)
#define CTA_SYNC() __syncthreads()
#define DYN_SMEM(type, name)                                           \
  extern __shared__ __align__(16) unsigned char tmf_dyn_smem_raw[];    \
  type *name = reinterpret_cast<type *>(tmf_dyn_smem_raw)

namespace tmf {
inline int check_cuda(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return TMF_OK;
  set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return TMF_ERR_RUNTIME;
}
// per-tag kernel timing (CUDA events around every launch) -- enabled only by bench.py's profiling
// pass through tmf_prof_enable(); the launch counter is always on (bench "gpu_launches").
void count_launch();
int ensure_max_dynamic_smem(const void *kernel);    // once per kernel: cudaFuncAttributeMaxDynamicSharedMemorySize = 227 KB
bool prof_enabled();
void prof_begin(const char *tag, void *stream);
void prof_end(void *stream);
template <class K, class... Args>
inline int launch_t(const char *tag, K kernel, int grid, int block, size_t smem_bytes, void *stream,
                    Args... args) {
  if (grid <= 0) return TMF_OK;
  if (smem_bytes > 48 * 1024) {
    // opt in to the full 227 KB once per kernel function (the attribute is per function, not per launch; the
    // call takes a context-wide lock, and with six pipeline threads launching it showed up as millisecond
    // stalls of unrelated launches)
    int rc = ensure_max_dynamic_smem(reinterpret_cast<const void *>(kernel));
    if (rc) return rc;
  }
  count_launch();
  const bool prof = prof_enabled();
  if (prof) prof_begin(tag, stream);
  static const bool dbg = std::getenv("TMF_DEBUG_LAUNCH") != nullptr;
  std::chrono::steady_clock::time_point t0;
  if (dbg) t0 = std::chrono::steady_clock::now();
  kernel<<<grid, block, smem_bytes, (cudaStream_t)stream>>>(args...);
  if (dbg) {
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    if (us > 40.0) std::fprintf(stderr, "[tmf launch] %-14s grid %6d block %4d smem %6zu : %8.1f us\n", tag, grid, block, smem_bytes, us);
  }
  if (prof) prof_end(stream);
  return check_cuda(cudaGetLastError(), "kernel launch");
}
template <class K, class... Args>
inline int launch(K kernel, int grid, int block, size_t smem_bytes, void *stream, Args... args) {
  return launch_t("other", kernel, grid, block, smem_bytes, stream, args...);
}
// Host -> device copy that never blocks the calling thread on the stream: pinned sources are copied
// directly; pageable sources (descriptor vectors) are first copied into a reusable pinned staging ring of
// the calling thread.  (cudaMemcpyAsync from pageable memory waits for all prior work of the stream -- with
// the event-gated pipeline that turned every descriptor upload into a wait for the previous chunks' kernels.)
int copy_h2d(void *dst, const void *src, size_t bytes, void *stream);
inline int copy_d2h_async(void *dst, const void *src, size_t bytes, void *stream) {
  if (bytes == 0) return TMF_OK;
  return check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream),
                    "cudaMemcpyAsync D2H");
}
inline int copy_d2h_sync(void *dst, const void *src, size_t bytes, void *stream) {
  if (bytes) {
    int rc = check_cuda(
        cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream),
        "cudaMemcpyAsync D2H");
    if (rc) return rc;
  }
  return check_cuda(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize");
}
inline int memset_dev(void *dst, int v, size_t bytes, void *stream) {
  if (bytes == 0) return TMF_OK;
  return check_cuda(cudaMemsetAsync(dst, v, bytes, (cudaStream_t)stream), "cudaMemsetAsync");
}
inline int stream_sync(void *stream) {
  return check_cuda(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize");
}
}  // namespace tmf
#endif

namespace tmf {
// Aligned bump allocator over a caller-provided workspace.
struct Arena {
  unsigned char *base;
  int64_t size, used = 0;
  Arena(void *p, int64_t bytes) : base(static_cast<unsigned char *>(p)), size(bytes) {}
  template <class T>
  T *take(int64_t count) {
    int64_t off = (used + 255) & ~int64_t(255);
    used = off + count * (int64_t)sizeof(T);
    return reinterpret_cast<T *>(base + off);  // validity is checked by the caller via ok()
  }
  bool ok() const { return base == nullptr || used <= size; }
};
inline int64_t align256(int64_t b) { return (b + 255) & ~int64_t(255); }
}  // namespace tmf
