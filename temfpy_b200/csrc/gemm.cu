// Grouped FP64 GEMM on the sm_100a DMMA pipe (mma.sync.m8n8k4.f64).
//
// One launch processes a list of independent, differently sized problems
//     C_j (M x N) = alpha * diag(row_scale) * op(A_j) * op(B_j) * diag(col_scale) + beta * C_j
// (all column-major).  It is the workhorse of the mode extraction (range finder B*Omega,
// projections Q^T B, Rayleigh-Ritz blocks), of the overlap O = V_bra^T V_ket (slater.py:1071),
// of C = Phi Phi^T (slater.py:1177) and of the Gutzwiller block products.
//
// CTA tile 64 x 64 x 16, 256 threads = 8 warps in a 4 x 2 arrangement, each warp owns a 16 x 32
// accumulator (2 x 4 DMMA tiles).  Operand tiles are staged in shared memory "k-contiguous" with a
// row stride of 20 doubles, which makes every fragment load (8 rows x 4 k) conflict-free, and are
// double-buffered with register prefetch of the next k-tile.  Tiles of all jobs are enumerated by
// a prefix table so that one grid covers the whole group (grid >> 148 for the chain workloads).
#include "cta.hpp"
#if !defined(TMF_HOSTSIM)
#include <cuda.h>
#include <map>
#include <mutex>
#include <tuple>
#endif

namespace tmf {

constexpr int TM = 64, TN = 64, TK = 16, LDS_STRIDE = 20;

struct GemmDesc {
  tmf_gemm_job job;
};
static_assert(sizeof(tmf_gemm_job) == 128, "descriptor must be 128 bytes");

TMF_DEVICE double load_a(const tmf_gemm_job &j, int m, int k) {
  if (m >= j.M || k >= j.K) return 0.0;
  if (j.transA) {
    int col = j.a_idx ? j.a_idx[m] : m;
    if (col < 0) return 0.0;
    return j.A[(int64_t)col * j.lda + k + j.a_row_off];
  }
  return j.A[(int64_t)k * j.lda + m];
}
TMF_DEVICE double load_b(const tmf_gemm_job &j, int k, int n) {
  if (n >= j.N || k >= j.K) return 0.0;
  if (j.transB) return j.B[(int64_t)k * j.ldb + n];
  int col = j.b_idx ? j.b_idx[n] : n;
  if (col < 0) return 0.0;
  return j.B[(int64_t)col * j.ldb + k + j.b_row_off];
}
TMF_DEVICE void store_c(const tmf_gemm_job &j, int m, int n, double acc) {
  if (m >= j.M || n >= j.N) return;
  double v = j.alpha * acc;
  if (j.row_scale) v *= j.row_scale[m];
  if (j.col_scale) v *= j.col_scale[n];
  double *p = j.C + (int64_t)n * j.ldc + m;
  if (j.beta != 0.0) v += j.beta * (*p);
  *p = v;
}

// binary search: largest j with prefix[j] <= tile
TMF_DEVICE int find_job(const int *prefix, int njobs, int tile) {
  int lo = 0, hi = njobs;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (prefix[mid] <= tile) lo = mid; else hi = mid;
  }
  return lo;
}

#if defined(TMF_HOSTSIM)
TMF_GLOBAL gemm_grouped_kernel(const tmf_gemm_job *jobs, const int *prefix, int njobs) {
  const int tile = BLOCK_ID;
  const int jid = find_job(prefix, njobs, tile);
  const tmf_gemm_job &j = jobs[jid];
  const int t = tile - prefix[jid];
  const int tiles_m = (j.M + TM - 1) / TM;
  const int m0 = (t % tiles_m) * TM, n0 = (t / tiles_m) * TN;
  for (int n = n0; n < n0 + TN && n < j.N; ++n)
    for (int m = m0; m < m0 + TM && m < j.M; ++m) {
      double acc = 0.0;
      for (int k = 0; k < j.K; ++k) acc += load_a(j, m, k) * load_b(j, k, n);
      store_c(j, m, n, acc);
    }
}
#else
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Shared-memory layouts of an operand tile (64 rows x 16 k):
//   KC  (k-contiguous in global): stored [row][k], row stride 20 doubles
//   MC  (row-contiguous in global): stored [k][row], k stride 72 doubles
// both make the DMMA fragment loads (8 rows x 4 k per warp) and the staging stores conflict-free.
constexpr int KC_STRIDE = 20, MC_STRIDE = 72;
constexpr int TILE_DOUBLES = 64 * KC_STRIDE;  // 1280 >= 16 * 72 = 1152

template <bool KC>
struct OperandLoader {
  const double *p[4];   // per-thread source pointers at k-tile 0 (nullptr: outside the matrix)
  int kk[4];            // k index inside the tile of each element
  int so[4];            // shared-memory offset of each element
  int64_t kstep;        // pointer increment per k-tile
  // rows: index along the tile's row dimension (M for A, N for B); `cols` maps it to a column of the
  // stored matrix when the operand is k-contiguous (optional gather), ld = leading dimension.
  __device__ void init(const double *base, int ld, const int *idx, int row0, int nrows, int roff, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      const int r = KC ? (e >> 4) : (e & 63);
      kk[i] = KC ? (e & 15) : (e >> 6);
      so[i] = KC ? r * KC_STRIDE + kk[i] : kk[i] * MC_STRIDE + r;
      const int g = row0 + r;
      p[i] = nullptr;
      if (g < nrows) {
        if (KC) {
          const int col = idx ? idx[g] : g;
          if (col >= 0) p[i] = base + (int64_t)col * ld + roff + kk[i];
        } else {
          p[i] = base + (int64_t)kk[i] * ld + g;
        }
      }
    }
    kstep = KC ? TK : (int64_t)TK * ld;
  }
  __device__ __forceinline__ void load(double (&r)[4], int k0, int K) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = (p[i] != nullptr && k0 + kk[i] < K) ? p[i][(int64_t)(k0 / TK) * kstep] : 0.0;
  }
  // the same with the k axis shifted by `sh` (0 / 1): tile position k' holds element k' - sh (zero at k' < sh)
  __device__ __forceinline__ void load_shifted(double (&r)[4], int k0, int K, int sh) const {
    const int64_t unit = kstep / TK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + kk[i] - sh;
      r[i] = (p[i] != nullptr && k >= 0 && k < K) ? p[i][(int64_t)(k0 / TK) * kstep - sh * unit] : 0.0;
    }
  }
  __device__ __forceinline__ void store(double *tile, const double (&r)[4]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) tile[so[i]] = r[i];
  }
  static __device__ __forceinline__ double frag(const double *tile, int row, int k) {
    return KC ? tile[row * KC_STRIDE + k] : tile[k * MC_STRIDE + row];
  }
};

template <bool AKC, bool BKC>
__device__ __forceinline__ void gemm_tile(const tmf_gemm_job &j, int m0, int n0, double *As, double *Bs) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  OperandLoader<AKC> la;
  OperandLoader<BKC> lb;
  // A: op(A) is M x K.  transA=1 -> stored K x M (k-contiguous, optional column gather a_idx)
  la.init(j.A, j.lda, j.a_idx, m0, j.M, j.a_row_off, tid);
  // B: op(B) is K x N.  transB=0 -> stored K x N (k-contiguous, optional column gather b_idx)
  lb.init(j.B, j.ldb, j.b_idx, n0, j.N, j.b_row_off, tid);
  double acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  const int nk = (j.K + TK - 1) / TK;
  double ra[4], rb[4];
  int buf = 0;
  if (nk > 0) {
    la.load(ra, 0, j.K);
    lb.load(rb, 0, j.K);
    la.store(As, ra);
    lb.store(Bs, rb);
  }
  __syncthreads();
  const int fr = lane >> 2, fk = lane & 3;
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) {
      la.load(ra, (kt + 1) * TK, j.K);
      lb.load(rb, (kt + 1) * TK, j.K);
    }
    const double *as = As + buf * TILE_DOUBLES, *bs = Bs + buf * TILE_DOUBLES;
#pragma unroll
    for (int ks = 0; ks < TK; ks += 4) {
      double fa[2], fb[4];
#pragma unroll
      for (int a = 0; a < 2; ++a) fa[a] = OperandLoader<AKC>::frag(as, wm + a * 8 + fr, ks + fk);
#pragma unroll
      for (int b = 0; b < 4; ++b) fb[b] = OperandLoader<BKC>::frag(bs, wn + b * 8 + fr, ks + fk);
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
    }
    if (kt + 1 < nk) {
      la.store(As + (buf ^ 1) * TILE_DOUBLES, ra);
      lb.store(Bs + (buf ^ 1) * TILE_DOUBLES, rb);
    }
    __syncthreads();
    buf ^= 1;
  }
  // epilogue: fragment (row = lane/4, cols = 2*(lane%4) + {0,1})
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int m = m0 + wm + a * 8 + fr, n = n0 + wn + b * 8 + 2 * fk;
      store_c(j, m, n, acc[a][b][0]);
      store_c(j, m, n + 1, acc[a][b][1]);
    }
}

// ---------------------------------------------------------------------------------------------
// TMA-staged variant for jobs whose A operand is a k-contiguous sub-block of the correlation matrix
// (W^T = B^T Q and A U0 of the mode extraction: ~85 % of the GEMM flops of a conversion).  One tensor map
// describes the whole L x L matrix; a job addresses its block by the coordinates of its origin (pad_[1] =
// first row, pad_[2] = first column), so no alignment condition falls on the block itself.  The A tiles
// (64 rows x 16 k = 8 KB) are fetched by cp.async.bulk.tensor into a 3-stage ring with the 128-byte swizzle
// (conflict-free fragment loads without padding) and signalled through mbarriers; one thread issues the
// copies two k-tiles ahead.  The (gathered / small) B operand keeps the register-prefetch path.  The
// per-iteration __syncthreads of the B double buffer also orders the reuse of an A stage, so no "empty"
// barriers are needed.  Rows / k beyond the block read neighbouring entries of C (finite) that meet zeros
// of the B tile or unused accumulator rows.
// ---------------------------------------------------------------------------------------------
constexpr int TMA_STAGES = 3;
constexpr int TMA_TILE_BYTES = 64 * TK * 8;   // 8192

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TMF_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TMF_MBAR_DONE;\n"
      "bra TMF_MBAR_WAIT;\n"
      "TMF_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

template <bool BKC>
__device__ __forceinline__ void gemm_tile_tma(const tmf_gemm_job &j, int m0, int n0, unsigned char *At, double *Bs,
                                              uint64_t *bars, const CUtensorMap *map) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  OperandLoader<BKC> lb;
  lb.init(j.B, j.ldb, j.b_idx, n0, j.N, j.b_row_off, tid);
  double acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  // The box of a bulk tensor copy must start on a 16-byte boundary, i.e. at an even column of C: a block that
  // starts at an odd column is read from the column before it (ksh = 1) and the k axis of the B tile is shifted
  // by one, with a zero in front (the extra column of C meets that zero).
  const int ksh = (j.pad_[2] + j.a_row_off) & 1;
  const int nk = (j.K + ksh + TK - 1) / TK;
  const int row0 = j.pad_[1] + m0, col0 = j.pad_[2] + j.a_row_off - ksh;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TMA_STAGES; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < TMA_STAGES - 1 && s < nk; ++s) {
      mbar_expect_tx(bars + s, TMA_TILE_BYTES);
      tma_load_2d(At + s * TMA_TILE_BYTES, map, col0 + s * TK, row0, bars + s);
    }
  }
  double rb[4];
  int buf = 0;
  if (nk > 0) {
    lb.load_shifted(rb, 0, j.K, ksh);
    lb.store(Bs, rb);
  }
  __syncthreads();
  const int fr = lane >> 2, fk = lane & 3;
  for (int kt = 0; kt < nk; ++kt) {
    // stage (kt + 2) % 3 was read last in iteration kt - 1, which every warp left through the barrier below
    if (tid == 0 && kt + TMA_STAGES - 1 < nk) {
      const int s = (kt + TMA_STAGES - 1) % TMA_STAGES;
      mbar_expect_tx(bars + s, TMA_TILE_BYTES);
      tma_load_2d(At + s * TMA_TILE_BYTES, map, col0 + (kt + TMA_STAGES - 1) * TK, row0, bars + s);
    }
    if (kt + 1 < nk) lb.load_shifted(rb, (kt + 1) * TK, j.K, ksh);
    const int st = kt % TMA_STAGES;
    mbar_wait(bars + st, (uint32_t)((kt / TMA_STAGES) & 1));
    const unsigned char *as = At + st * TMA_TILE_BYTES;
    const double *bs = Bs + buf * TILE_DOUBLES;
#pragma unroll
    for (int ks = 0; ks < TK; ks += 4) {
      double fa[2], fb[4];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        // 128-byte swizzle: the 16-byte chunk index of an element is XORed with (row & 7)
        const int r = wm + a * 8 + fr, k = ks + fk;
        fa[a] = *reinterpret_cast<const double *>(as + r * 128 + ((((k >> 1) ^ (r & 7))) << 4) + ((k & 1) << 3));
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) fb[b] = OperandLoader<BKC>::frag(bs, wn + b * 8 + fr, ks + fk);
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
    }
    if (kt + 1 < nk) lb.store(Bs + (buf ^ 1) * TILE_DOUBLES, rb);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int m = m0 + wm + a * 8 + fr, n = n0 + wn + b * 8 + 2 * fk;
      store_c(j, m, n, acc[a][b][0]);
      store_c(j, m, n + 1, acc[a][b][1]);
    }
}

__global__ void __launch_bounds__(256, 2)
gemm_grouped_tma_kernel(const tmf_gemm_job *__restrict__ jobs, const int *__restrict__ prefix, int njobs,
                        const __grid_constant__ CUtensorMap tmapC) {
  __shared__ __align__(1024) unsigned char At[TMA_STAGES * TMA_TILE_BYTES];   // 24 KB; doubles as As of the plain path
  static_assert(TMA_STAGES * TMA_TILE_BYTES >= 2 * TILE_DOUBLES * 8, "A staging must cover the plain double buffer");
  double *As = reinterpret_cast<double *>(At);
  __shared__ double Bs[2 * TILE_DOUBLES];
  __shared__ __align__(8) uint64_t bars[TMA_STAGES];
  __shared__ tmf_gemm_job js;
  const int tile = blockIdx.x;
  const int jid = find_job(prefix, njobs, tile);
  if (threadIdx.x < sizeof(tmf_gemm_job) / 8)
    reinterpret_cast<uint64_t *>(&js)[threadIdx.x] =
        reinterpret_cast<const uint64_t *>(&jobs[jid])[threadIdx.x];
  __syncthreads();
  const tmf_gemm_job &j = js;
  const int t = tile - prefix[jid];
  const int tiles_m = (j.M + TM - 1) / TM;
  const int m0 = (t % tiles_m) * TM, n0 = (t / tiles_m) * TN;
  const bool akc = j.transA != 0, bkc = j.transB == 0;
  if (j.pad_[0] == 1 && akc && j.a_idx == nullptr) {
    if (bkc) gemm_tile_tma<true>(j, m0, n0, At, Bs, bars, &tmapC);
    else gemm_tile_tma<false>(j, m0, n0, At, Bs, bars, &tmapC);
    return;
  }
  if (akc) {
    if (bkc) gemm_tile<true, true>(j, m0, n0, As, Bs);
    else gemm_tile<true, false>(j, m0, n0, As, Bs);
  } else {
    if (bkc) gemm_tile<false, true>(j, m0, n0, As, Bs);
    else gemm_tile<false, false>(j, m0, n0, As, Bs);
  }
}

__global__ void __launch_bounds__(256, 2)
gemm_grouped_kernel(const tmf_gemm_job *__restrict__ jobs, const int *__restrict__ prefix,
                    int njobs) {
  __shared__ double As[2 * TILE_DOUBLES];
  __shared__ double Bs[2 * TILE_DOUBLES];
  __shared__ tmf_gemm_job js;
  const int tile = blockIdx.x;
  const int jid = find_job(prefix, njobs, tile);
  if (threadIdx.x < sizeof(tmf_gemm_job) / 8)
    reinterpret_cast<uint64_t *>(&js)[threadIdx.x] =
        reinterpret_cast<const uint64_t *>(&jobs[jid])[threadIdx.x];
  __syncthreads();
  const tmf_gemm_job &j = js;
  const int t = tile - prefix[jid];
  const int tiles_m = (j.M + TM - 1) / TM;
  const int m0 = (t % tiles_m) * TM, n0 = (t / tiles_m) * TN;
  const bool akc = j.transA != 0, bkc = j.transB == 0;
  if (akc) {
    if (bkc) gemm_tile<true, true>(j, m0, n0, As, Bs);
    else gemm_tile<true, false>(j, m0, n0, As, Bs);
  } else {
    if (bkc) gemm_tile<false, true>(j, m0, n0, As, Bs);
    else gemm_tile<false, false>(j, m0, n0, As, Bs);
  }
}
#endif

// Host launcher: builds the tile prefix table and uploads descriptors + table into desc_dev.
int gemm_grouped(const tmf_gemm_job *jobs, int njobs, void *desc_dev, void *stream) {
  if (njobs <= 0) return TMF_OK;
  std::vector<int> prefix(njobs + 1, 0);
  for (int i = 0; i < njobs; ++i) {
    int tm = (jobs[i].M + TM - 1) / TM, tn = (jobs[i].N + TN - 1) / TN;
    if (jobs[i].M <= 0 || jobs[i].N <= 0) tm = tn = 0;
    prefix[i + 1] = prefix[i] + tm * tn;
  }
  const int ntiles = prefix[njobs];
  if (ntiles == 0) return TMF_OK;
  unsigned char *d = static_cast<unsigned char *>(desc_dev);
  int rc = copy_h2d(d, jobs, sizeof(tmf_gemm_job) * (size_t)njobs, stream);
  if (rc) return rc;
  int *dprefix = reinterpret_cast<int *>(d + sizeof(tmf_gemm_job) * (size_t)njobs);
  rc = copy_h2d(dprefix, prefix.data(), sizeof(int) * (size_t)(njobs + 1), stream);
  if (rc) return rc;
  return launch_t("gemm", gemm_grouped_kernel, ntiles, 256, 0, stream,
                  reinterpret_cast<const tmf_gemm_job *>(d), (const int *)dprefix, njobs);
}

int gemm_launch_uploaded(const tmf_gemm_job *jobs_dev, const int *prefix_dev, int njobs, int ntiles,
                         void *stream, const char *tag) {
  if (ntiles <= 0) return TMF_OK;
  return launch_t(tag, gemm_grouped_kernel, ntiles, 256, 0, stream, jobs_dev, prefix_dev, njobs);
}

// Launch with the A operands flagged pad_[0] = 1 staged by TMA from the matrix `Cmat` (L x L doubles, row pitch
// ldc).  Falls back to the plain kernel when the matrix does not meet the tensor-map conditions (16-byte
// aligned base and pitch) or the driver entry point is missing.
int gemm_launch_uploaded_tma(const tmf_gemm_job *jobs_dev, const int *prefix_dev, int njobs, int ntiles,
                             void *stream, const char *tag, const double *Cmat, int L, int ldc) {
  if (ntiles <= 0) return TMF_OK;
#if !defined(TMF_HOSTSIM)
  static const bool off = std::getenv("TMF_NO_TMA") != nullptr;
  if (!off && Cmat != nullptr && (reinterpret_cast<uintptr_t>(Cmat) & 15) == 0 && (ldc & 1) == 0 && L >= TK) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
          qres != cudaDriverEntryPointSuccess)
        fn = nullptr;
      return reinterpret_cast<EncodeFn>(fn);
    }();
    if (encode != nullptr) {
      static std::mutex mu;
      static std::map<std::tuple<const void *, int, int>, CUtensorMap> cache;
      CUtensorMap map;
      bool ok = true;
      {
        std::lock_guard<std::mutex> lk(mu);
        auto key = std::make_tuple((const void *)Cmat, L, ldc);
        auto it = cache.find(key);
        if (it == cache.end()) {
          if (cache.size() > 64) cache.clear();
          const cuuint64_t dims[2] = {(cuuint64_t)L, (cuuint64_t)L};
          const cuuint64_t strides[1] = {(cuuint64_t)ldc * 8};
          const cuuint32_t box[2] = {(cuuint32_t)TK, 64};
          const cuuint32_t estr[2] = {1, 1};
          ok = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(Cmat), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
          if (ok) cache[key] = map;
        } else {
          map = it->second;
        }
      }
      if (ok) return launch_t(tag, gemm_grouped_tma_kernel, ntiles, 256, 0, stream, jobs_dev, prefix_dev, njobs, map);
    }
  }
#else
  (void)Cmat; (void)L; (void)ldc;
#endif
  return gemm_launch_uploaded(jobs_dev, prefix_dev, njobs, ntiles, stream, tag);
}

int64_t gemm_desc_bytes(int njobs) {
  return align256((int64_t)sizeof(tmf_gemm_job) * njobs + 4 * (int64_t)(njobs + 1));
}

}  // namespace tmf

extern "C" int64_t tmf_gemm_desc_bytes(int njobs) { return tmf::gemm_desc_bytes(njobs) + 256; }

extern "C" int tmf_gemm_grouped(const tmf_gemm_job *jobs_host, int njobs, void *desc_dev,
                                void *stream) {
  return tmf::gemm_grouped(jobs_host, njobs, desc_dev, stream);
}

extern "C" int tmf_corr_build(const double *phi_dev, int L, int N, int ldphi, double *C_dev,
                              int ldc, void *stream) {
  // slater.py:1177: C = Phi Phi^T with Phi given row-major (L x N, row stride ldphi), i.e. as the
  // column-major N x L matrix M = Phi^T: C = M^T M, both operands k-contiguous.
  static void *scratch = nullptr;
#if !defined(TMF_HOSTSIM)
  if (!scratch) {
    if (cudaMalloc(&scratch, 1024) != cudaSuccess) {
      tmf::set_error("cudaMalloc(descriptor scratch) failed");
      return TMF_ERR_RUNTIME;
    }
  }
#else
  static unsigned char host_scratch[1024];
  scratch = host_scratch;
#endif
  tmf_gemm_job j;
  std::memset(&j, 0, sizeof(j));
  j.A = phi_dev; j.B = phi_dev; j.C = C_dev;
  j.M = L; j.N = L; j.K = N;
  j.lda = ldphi; j.ldb = ldphi; j.ldc = ldc;
  j.transA = 1; j.transB = 0;
  j.alpha = 1.0; j.beta = 0.0;
  return tmf::gemm_grouped(&j, 1, scratch, stream);
}

// ---------------------------------------------------------------------------------------------
// Projector test of the input (slater.C_to_MPS): max |(C C - C)[i, j]| over the first `rows` rows, C symmetric
// L x L with pitch ldc.  The product runs on the grouped GEMM, the comparison in one CTA; `work_dev` holds
// rows * L + 1 doubles, the result is the last of them.
// ---------------------------------------------------------------------------------------------
namespace tmf {
TMF_GLOBAL defect_kernel(const double *T, const double *C, int rows, int L, int ldc, double *out) {
  DYN_SMEM(double, part);
  PAR_FOR(t, NTHREADS) {
    double m = 0.0;
    for (int64_t idx = t; idx < (int64_t)rows * L; idx += NTHREADS) {
      const int j = (int)(idx / rows), i = (int)(idx - (int64_t)j * rows);
      const double d = fabs(T[idx] - C[(int64_t)i * ldc + j]);
      m = (d > m || d != d) ? d : m;          // (a NaN wins)
    }
    part[t] = m;
  }
  CTA_SYNC();
  PAR_FOR(one, 1) {
    double m = 0.0;
    for (int t = 0; t < NTHREADS; ++t) m = (part[t] > m || part[t] != part[t]) ? part[t] : m;
    out[0] = m;
  }
}
}  // namespace tmf

extern "C" int tmf_projector_defect(const double *C_dev, int L, int ldc, int rows, double *work_dev, void *desc_dev,
                                    void *stream) {
  using namespace tmf;
  if (rows > L) rows = L;
  tmf_gemm_job j;
  std::memset(&j, 0, sizeof(j));
  // T (rows x L, column-major) = C[:, :rows]^T C = (C C)[:rows, :]
  j.A = C_dev; j.lda = ldc; j.transA = 1;
  j.B = C_dev; j.ldb = ldc; j.transB = 0;
  j.C = work_dev; j.ldc = rows;
  j.M = rows; j.N = L; j.K = L;
  j.alpha = 1.0; j.beta = 0.0;
  int rc = gemm_grouped(&j, 1, desc_dev, stream);
  if (rc) return rc;
  return launch_t("defect", defect_kernel, 1, 512, 512 * sizeof(double), stream, (const double *)work_dev, C_dev, rows,
                  L, ldc, work_dev + (int64_t)rows * L);
}
