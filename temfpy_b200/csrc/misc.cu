// Library identity, device queries and the FP64 peak probe used as roofline denominator.
#include "cta.hpp"

#if defined(TMF_HOSTSIM)
namespace tmfsim {
thread_local int block_id = 0;
thread_local int n_threads = 1;
thread_local unsigned char *smem = nullptr;
}  // namespace tmfsim
#endif

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>

namespace tmf {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
#if !defined(TMF_HOSTSIM)
static bool g_prof = false;
static std::mutex g_prof_mu;
struct ProfRec { std::string tag; cudaEvent_t a, b; void *stream; };
static thread_local size_t t_last = 0;   // launches are issued from several pipeline threads
static std::vector<ProfRec> g_recs;
#if !defined(TMF_HOSTSIM)
namespace {
struct StagingRing {
  unsigned char *p = nullptr;
  size_t cap = 0, off = 0;
  cudaEvent_t ev = nullptr;
  bool pending = false;
  std::mutex mu;
  ~StagingRing() {
    if (ev) cudaEventDestroy(ev);
    if (p) cudaFreeHost(p);
  }
  // returns pinned space for `bytes` (nullptr: use the pageable path)
  unsigned char *take(size_t bytes) {
    const size_t need = (bytes + 255) & ~size_t(255);
    if (need > (size_t(8) << 20)) return nullptr;
    if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (off + need > cap) {
      if (pending) { cudaEventSynchronize(ev); pending = false; }   // every earlier copy from the ring is done
      off = 0;
      if (need > cap) {
        if (p) cudaFreeHost(p);
        cap = std::max(need * 2, size_t(4) << 20);
        if (cudaHostAlloc(reinterpret_cast<void **>(&p), cap, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; cap = 0; return nullptr; }
      }
    }
    unsigned char *r = p + off;
    off += need;
    return r;
  }
};
// one ring per stream (the pipeline's streams are persistent; its worker threads are not), never freed
std::mutex g_ring_mu;
std::map<void *, StagingRing *> g_rings;
StagingRing *ring_of(void *stream) {
  std::lock_guard<std::mutex> lk(g_ring_mu);
  StagingRing *&r = g_rings[stream];
  if (!r) r = new StagingRing();
  return r;
}
}  // namespace

int copy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
  if (bytes == 0) return TMF_OK;
  cudaPointerAttributes attr;
  const bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  if (!pinned) {
    cudaGetLastError();   // (older drivers flag unregistered pointers as an error)
    StagingRing *ring = ring_of(stream);
    std::lock_guard<std::mutex> lk(ring->mu);
    if (unsigned char *st = ring->take(bytes)) {
      std::memcpy(st, src, bytes);
      int rc = check_cuda(cudaMemcpyAsync(dst, st, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), "cudaMemcpyAsync H2D");
      if (rc) return rc;
      cudaEventRecord(ring->ev, (cudaStream_t)stream);
      ring->pending = true;
      return TMF_OK;
    }
  }
  return check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), "cudaMemcpyAsync H2D");
}
#endif

#if !defined(TMF_HOSTSIM)
int ensure_max_dynamic_smem(const void *kernel) {
  static std::mutex mu;
  static std::vector<const void *> done;
  std::lock_guard<std::mutex> lk(mu);
  for (const void *k : done)
    if (k == kernel) return TMF_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute");
  done.push_back(kernel);
  return TMF_OK;
}
#endif
bool prof_enabled() { return g_prof; }
void prof_begin(const char *tag, void *stream) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.tag = tag;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  r.stream = stream;
  cudaEventRecord(r.a, (cudaStream_t)stream);
  t_last = g_recs.size();
  g_recs.push_back(r);
}
void prof_end(void *stream) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_recs[t_last].b, (cudaStream_t)stream);
}
#endif
}  // namespace tmf

extern "C" long long tmf_launch_count(int reset) {
  long long v = tmf::g_launches.load();
  if (reset) tmf::g_launches.store(0);
  return v;
}

extern "C" int tmf_prof_enable(int on) {
#if !defined(TMF_HOSTSIM)
  std::lock_guard<std::mutex> lk(tmf::g_prof_mu);
  for (auto &r : tmf::g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  tmf::g_recs.clear();
  tmf::g_prof = on != 0;
#else
  (void)on;
#endif
  return TMF_OK;
}

// writes "tag total_ms launches\n" lines into buf (synchronises the device)
extern "C" int tmf_prof_report(char *buf, int cap) {
  std::string out;
#if !defined(TMF_HOSTSIM)
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(tmf::g_prof_mu);
  std::map<std::string, std::pair<double, int>> acc;
  for (auto &r : tmf::g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      acc[r.tag].first += ms;
      acc[r.tag].second += 1;
    }
  }
  for (auto &kv : acc)
    out += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
#endif
  if ((int)out.size() + 1 > cap) return TMF_ERR_VALUE;
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return TMF_OK;
}

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

// timeline of the recorded launches: "tag stream start_ms end_ms" per line, relative to the first
// recorded launch (profiling pass of bench.py / scratch tools; synchronises the device)
extern "C" int tmf_prof_timeline(char *buf, int cap) {
  std::string out;
#if !defined(TMF_HOSTSIM)
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(tmf::g_prof_mu);
  if (!tmf::g_recs.empty()) {
    std::map<void *, int> sid;
    cudaEvent_t base = tmf::g_recs[0].a;
    for (auto &r : tmf::g_recs) {
      float t0 = 0.f, t1 = 0.f;
      if (cudaEventElapsedTime(&t0, base, r.a) != cudaSuccess) continue;
      if (cudaEventElapsedTime(&t1, base, r.b) != cudaSuccess) continue;
      if (!sid.count(r.stream)) { int n = (int)sid.size(); sid[r.stream] = n; }
      out += r.tag + " " + std::to_string(sid[r.stream]) + " " + std::to_string(t0) + " " + std::to_string(t1) + "\n";
    }
  }
#endif
  if ((int)out.size() + 1 > cap) return TMF_ERR_VALUE;
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return TMF_OK;
}

// ---------------------------------------------------------------------------------------------
// Peer windows (multi-GPU gather fused into the tensor kernels) and shared pinned host segments.
//
// A *peer window* is a buffer in the HBM of the destination rank that the other ranks of the node map through
// CUDA IPC; their minors kernels then store the site tensors straight into it over NVLink (P2P stores), so the
// "gather" of a sharded conversion is the kernels' own output traffic.  The host segments are POSIX shared memory
// mapped by every rank and registered with the driver, so that each GPU copies its shard to the destination
// process's address space over its own PCIe link.
// ---------------------------------------------------------------------------------------------
extern "C" int tmf_ipc_export(const void *dev_ptr, unsigned char *handle64, int64_t *offset_out) {
#if defined(TMF_HOSTSIM)
  (void)dev_ptr; (void)handle64; (void)offset_out;
  tmf::set_error("peer windows need the CUDA build");
  return TMF_ERR_RUNTIME;
#else
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  typedef int (*RangeFn)(unsigned long long *, size_t *, unsigned long long);
  static RangeFn range = [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<RangeFn>(fn);
  }();
  if (!range) { tmf::set_error("cuMemGetAddressRange unavailable"); return TMF_ERR_RUNTIME; }
  unsigned long long base = 0;
  size_t size = 0;
  if (range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0) {
    tmf::set_error("cuMemGetAddressRange failed");
    return TMF_ERR_RUNTIME;
  }
  cudaIpcMemHandle_t h;
  int rc = tmf::check_cuda(cudaIpcGetMemHandle(&h, reinterpret_cast<void *>((uintptr_t)base)), "cudaIpcGetMemHandle");
  if (rc) return rc;
  std::memcpy(handle64, &h, 64);
  *offset_out = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
  return TMF_OK;
#endif
}

extern "C" int tmf_ipc_open(const unsigned char *handle64, void **base_out) {
#if defined(TMF_HOSTSIM)
  (void)handle64; (void)base_out;
  tmf::set_error("peer windows need the CUDA build");
  return TMF_ERR_RUNTIME;
#else
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  // (the flag enables peer access from the current device to the owner of the allocation)
  return tmf::check_cuda(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
#endif
}

extern "C" int tmf_ipc_close(void *base) {
#if defined(TMF_HOSTSIM)
  (void)base;
  return TMF_OK;
#else
  return tmf::check_cuda(cudaIpcCloseMemHandle(base), "cudaIpcCloseMemHandle");
#endif
}

extern "C" int tmf_host_register(void *ptr, int64_t bytes) {
#if defined(TMF_HOSTSIM)
  (void)ptr; (void)bytes;
  return TMF_OK;
#else
  return tmf::check_cuda(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable), "cudaHostRegister");
#endif
}

extern "C" int tmf_host_unregister(void *ptr) {
#if defined(TMF_HOSTSIM)
  (void)ptr;
  return TMF_OK;
#else
  return tmf::check_cuda(cudaHostUnregister(ptr), "cudaHostUnregister");
#endif
}

extern "C" int tmf_copy_d2h_async(void *dst_host, const void *src_dev, int64_t bytes, void *stream) {
  return tmf::copy_d2h_async(dst_host, src_dev, (size_t)bytes, stream);
}
