// K10: all minors det(S[rows(alpha)][:, cols(beta)]) of every charge block of every site.
//
// reference: slater.py:828-869 (_tensor_block: gather an (nsb, nsk, n, n) array and call a batched
// LAPACK det, one LU of size n ~ 11 per tensor entry; 87.5 % of the reference's run time) and the
// det_always scaling of :1137.
//
// Instead of one n x n LU per entry we share the elimination between all kets of a bra row:
//   X = S[rows(alpha), :]   (n x s_ket)  is row-reduced once (Gauss-Jordan, pivoting along each row)
//   to [I | Y] on a pivot column set C0(alpha);  for any column set C with |C| = n
//      det X[:, C] = (prod of pivots) * sign * det Y[C0 \ C, C \ C0],
//   a determinant of size d = |C \ C0| (d <= 5 at chi = 1024, mean 2.5, versus n = 11).
// The pivot of row t is its largest entry among the unused columns (partial pivoting along the row).
//
// One CTA handles up to MB_ROWS bra rows of one block.  Its warps work independently: a warp row-reduces
// one bra row in a private shared-memory tile (lanes = columns, __syncwarp only), then every lane
// evaluates one (alpha, beta) entry with the d x d matrix in registers; stores are coalesced along beta.
#include "cta.hpp"

namespace tmf {

constexpr int MB_ROWS = 32;    // bra rows per CTA
constexpr int DGEN = 16;       // generic local-memory path up to 16 x 16
static_assert(sizeof(tmf_minor_block) == 64, "block descriptor must be 64 bytes");

// generic fallback (d > DMAX): in-place LU on a local array
#if !defined(TMF_HOSTSIM)
__device__ __noinline__
#else
static
#endif
double det_generic(double *m, int d, int ld) {
  double det = 1.0;
#pragma unroll 1
  for (int j = 0; j < d; ++j) {
    int p = j;
    double best = fabs(m[j * ld + j]);
#pragma unroll 1
    for (int i = j + 1; i < d; ++i)
      if (fabs(m[i * ld + j]) > best) { best = fabs(m[i * ld + j]); p = i; }
    if (best == 0.0) return 0.0;
    if (p != j) {
      det = -det;
#pragma unroll 1
      for (int c = 0; c < d; ++c) { double t = m[j * ld + c]; m[j * ld + c] = m[p * ld + c]; m[p * ld + c] = t; }
    }
    double piv = m[j * ld + j];
    det *= piv;
#pragma unroll 1
    for (int i = j + 1; i < d; ++i) {
      double l = m[i * ld + j] / piv;
#pragma unroll 1
      for (int c = j + 1; c < d; ++c) m[i * ld + c] -= l * m[j * ld + c];
    }
  }
  return det;
}

TMF_DEVICE int popc64(uint64_t x) {
#if defined(TMF_HOSTSIM)
  return __builtin_popcountll(x);
#else
  return __popcll(x);
#endif
}
TMF_DEVICE int ctz64(uint64_t x) {
#if defined(TMF_HOSTSIM)
  return __builtin_ctzll(x);
#else
  return __ffsll((long long)x) - 1;
#endif
}

// entry with more than 4 exchanged columns (rare): LU on a local-memory array
#if !defined(TMF_HOSTSIM)
__device__ __noinline__
#else
static
#endif
double entry_generic(const double *x, int smax, const int *colrow, uint64_t U, uint64_t mu,
                                uint64_t de, int d, int &par) {
  if (d > DGEN) return NAN;   // would need > 16 simultaneous column exchanges
  double buf[DGEN * DGEN];
  int rows_[DGEN], cols_[DGEN];
#pragma unroll 1
  for (int i = 0; i < d; ++i) {
    const int m = ctz64(mu), e = ctz64(de);
    mu &= mu - 1;
    de &= de - 1;
    const int a = m < e ? m : e, b = m < e ? e : m;
    const uint64_t between = (b - a > 1) ? (((1ull << (b - a - 1)) - 1) << (a + 1)) : 0ull;
    par += popc64(U & between);
    rows_[i] = colrow[m];
    cols_[i] = e;
  }
#pragma unroll 1
  for (int i = 0; i < d; ++i)
#pragma unroll 1
    for (int j = 0; j < d; ++j) buf[i * DGEN + j] = x[rows_[i] * smax + cols_[j]];
  return det_generic(buf, d, DGEN);
}

// Closed-form determinants for the reduced sizes that occur in practice (d <= 4 covers > 99.9 % of
// the entries at chi = 1024).  Straight-line cofactor / 2x2-minor expansions keep the kernel small
// enough for the instruction cache (an unrolled pivoted LU per size made it 340 KB of SASS and
// stalled the warps on instruction fetch).
TMF_DEVICE double det2(double a, double b, double c, double d) { return a * d - b * c; }
TMF_DEVICE double det3(const double *m) {   // row-major 3 x 3
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
         m[2] * (m[3] * m[7] - m[4] * m[6]);
}
TMF_DEVICE double det4(const double *m) {   // row-major 4 x 4, complementary 2 x 2 minors of rows 01 | 23
  const double s0 = m[0] * m[5] - m[1] * m[4], s1 = m[0] * m[6] - m[2] * m[4], s2 = m[0] * m[7] - m[3] * m[4];
  const double s3 = m[1] * m[6] - m[2] * m[5], s4 = m[1] * m[7] - m[3] * m[5], s5 = m[2] * m[7] - m[3] * m[6];
  const double c5 = m[10] * m[15] - m[11] * m[14], c4 = m[9] * m[15] - m[11] * m[13], c3 = m[9] * m[14] - m[10] * m[13];
  const double c2 = m[8] * m[15] - m[11] * m[12], c1 = m[8] * m[14] - m[10] * m[12], c0 = m[8] * m[13] - m[9] * m[12];
  return s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
}

// One tensor entry: gathers Y[C0 \\ C, C \\ C0] (d x d) from the reduced matrix, accumulates the
// permutation parity and returns the determinant.
TMF_DEVICE double entry_value(const double *x, int smax, const int *colrow, uint64_t c0, uint64_t cm) {
  const uint64_t U = cm & c0;
  uint64_t mu = c0 & ~cm, de = cm & ~c0;
  const int d = popc64(de);
  if (d == 0) return 1.0;
  int par = 0;
  double val;
  if (d <= 4) {
    int rr[4] = {0, 0, 0, 0}, cc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < d) {
        const int m = ctz64(mu), e = ctz64(de);
        mu &= mu - 1;
        de &= de - 1;
        const int a = m < e ? m : e, b = m < e ? e : m;
        const uint64_t between = (b - a > 1) ? (((1ull << (b - a - 1)) - 1) << (a + 1)) : 0ull;
        par += popc64(U & between);
        rr[i] = colrow[m] * smax;
        cc[i] = e;
      }
    }
    if (d == 1) {
      val = x[rr[0] + cc[0]];
    } else if (d == 2) {
      val = det2(x[rr[0] + cc[0]], x[rr[0] + cc[1]], x[rr[1] + cc[0]], x[rr[1] + cc[1]]);
    } else if (d == 3) {
      double m[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) m[3 * i + j] = x[rr[i] + cc[j]];
      val = det3(m);
    } else {
      double m[16];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) m[4 * i + j] = x[rr[i] + cc[j]];
      val = det4(m);
    }
  } else {
    val = entry_generic(x, smax, colrow, U, mu, de, d, par);
  }
  return (par & 1) ? -val : val;
}

// ---------------------------------------------------------------------------------------------
// Warp-independent kernel: after the sometimes matrix is staged in shared memory (one CTA barrier)
// every warp processes bra rows on its own -- row reduction with lanes = columns in a private
// shared-memory tile, then one tensor entry per lane -- synchronising with __syncwarp only.
// ---------------------------------------------------------------------------------------------
#if defined(TMF_HOSTSIM)
#define LANE_FOR(l) for (int l = 0; l < 32; ++l)
#define WARP_FOR(w, W) for (int w = 0; w < (W); ++w)
#define WSYNC() ((void)0)
#else
#define LANE_FOR(l) for (int l = (threadIdx.x & 31), l##_once = 1; l##_once; l##_once = 0)
#define WARP_FOR(w, W) for (int w = (threadIdx.x >> 5), w##_once = 1; w##_once && w < (W); w##_once = 0)
#define WSYNC() __syncwarp()
#endif

struct WarpMeta {
  uint64_t c0;        // pivot column set
  double scale;       // prod of pivots * sigma0 * det_always
  double inv;         // 1 / current pivot
  int pc, pad_;       // current pivot column
  int colrow[64];     // pivot row of every pivot column
  double cand[64];    // simulator only: candidates of the pivot search
};

constexpr int MB_WARPS = 8;

TMF_GLOBAL minors_kernel(const tmf_minor_block *blocks, const int *cta_prefix, int nblocks,
                         int nmax, int smax) {
  int lo = 0, hi = nblocks;
  const int cta = BLOCK_ID;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cta_prefix[mid] <= cta) lo = mid; else hi = mid;
  }
  const tmf_minor_block blk = blocks[lo];
  const int row0 = (cta - cta_prefix[lo]) * MB_ROWS;
  const int nrows = (blk.n_bra - row0 < MB_ROWS) ? (blk.n_bra - row0) : MB_ROWS;
  const int n = blk.minor, sk = blk.s_ket, sb = blk.s_bra;

  DYN_SMEM(unsigned char, raw);
  double *Ssm = reinterpret_cast<double *>(raw);                 // smax * smax
  double *Xall = Ssm + (size_t)smax * smax;                      // MB_WARPS * nmax * smax
  WarpMeta *metas = reinterpret_cast<WarpMeta *>(Xall + (size_t)MB_WARPS * nmax * smax);

  const double det_always = blk.det ? *blk.det : 1.0;
  PAR_FOR(idx, sb * sk) Ssm[idx] = blk.S[idx];
  CTA_SYNC();

  if (n == 0) {  // empty minors: det of a 0 x 0 matrix is 1
    PAR_FOR(idx, nrows * blk.n_ket) {
      int a = idx / blk.n_ket, c = idx - a * blk.n_ket;
      blk.out[(int64_t)(row0 + a) * blk.n_ket + c] = det_always;
    }
    return;
  }

  WARP_FOR(w, MB_WARPS) {
    double *x = Xall + (size_t)w * nmax * smax;
    WarpMeta *mt = metas + w;
    for (int a = w; a < nrows; a += MB_WARPS) {
      const uint64_t rmask = blk.bra_masks[row0 + a];
      // ---- gather X = S[rows(alpha), :] (lane = column) -------------------------------------
      LANE_FOR(l) {
        for (int c = l; c < sk; c += 32) {
          uint64_t rm = rmask;
          int r = 0;
          while (rm) {
            const int bit = ctz64(rm);
            rm &= rm - 1;
            x[r * smax + c] = Ssm[(size_t)c * sb + bit];
            ++r;
          }
        }
        if (l == 0) { mt->c0 = 0; mt->scale = det_always; }
      }
      WSYNC();
      // ---- Gauss-Jordan, pivoting along the row (largest entry among the unused columns) ------
      for (int t = 0; t < n; ++t) {
#if defined(TMF_HOSTSIM)
        LANE_FOR(l) {
          for (int c = l; c < sk; c += 32)
            mt->cand[c] = ((mt->c0 >> c) & 1) ? -1.0 : fabs(x[t * smax + c]);
        }
        LANE_FOR(l) if (l == 0) {
          double best = -1.0;
          int bc = 0;
          for (int c = 0; c < sk; ++c)
            if (mt->cand[c] > best) { best = mt->cand[c]; bc = c; }
          const double pv = x[t * smax + bc];
          mt->pc = bc;
          mt->inv = (pv != 0.0) ? 1.0 / pv : 0.0;
          mt->scale *= pv;
          mt->c0 |= (1ull << bc);
          mt->colrow[bc] = t;
        }
#else
        {
          // pivot = largest |x[t][c]| among the unused columns.  The magnitudes are compared as
          // float bit patterns (monotone for non-negative values) with the column index packed in
          // the low 6 bits, so that one 32-bit warp reduction (REDUX) replaces five shuffle rounds;
          // 18 bits of mantissa are ample for choosing a pivot.
          const int l = threadIdx.x & 31;
          const uint64_t c0 = mt->c0;
          unsigned key = 0u;
          for (int c = l; c < sk; c += 32) {
            if ((c0 >> c) & 1) continue;
            const unsigned k2 = (__float_as_uint(fabsf((float)x[t * smax + c])) & ~63u) | (unsigned)(63 - c);
            key = k2 > key ? k2 : key;
          }
          key = __reduce_max_sync(0xffffffffu, key);
          if (l == 0) {
            int bc = 63 - (int)(key & 63u);
            if (key == 0u) {   // every unused entry is zero (or no column left): singular minor
              bc = 0;
              while (bc < sk - 1 && ((c0 >> bc) & 1)) ++bc;
            }
            const double pv = x[t * smax + bc];
            mt->pc = bc;
            mt->inv = (pv != 0.0) ? 1.0 / pv : 0.0;
            mt->scale *= pv;
            mt->c0 = c0 | (1ull << bc);
            mt->colrow[bc] = t;
          }
        }
#endif
        WSYNC();
        LANE_FOR(l) {
          const int pc = mt->pc;
          const double inv = mt->inv;
          for (int c = l; c < sk; c += 32) {
            if (c == pc) continue;
            const double u = x[t * smax + c] * inv;
            x[t * smax + c] = u;
#pragma unroll 4
            for (int r = 0; r < n; ++r)
              if (r != t) x[r * smax + c] -= x[r * smax + pc] * u;
          }
        }
        WSYNC();
        // (pivot columns are never read again: later pivots are searched among the unused columns and
        //  the entries only gather columns outside the pivot set, so they are not reset to unit vectors)
      }
      // sigma0: sign of the permutation (rank of pivot column) -> pivot row
      LANE_FOR(l) if (l == 0) {
        uint64_t c0 = mt->c0, seen = 0;
        int inv = 0;
        while (c0) {
          const int c = ctz64(c0);
          c0 &= c0 - 1;
          const int r = mt->colrow[c];
          inv += popc64(seen >> (r + 1));
          seen |= (1ull << r);
        }
        if (inv & 1) mt->scale = -mt->scale;
      }
      WSYNC();
      // ---- one tensor entry per lane ---------------------------------------------------------
      LANE_FOR(l) {
        const uint64_t c0 = mt->c0;
        const double scale = mt->scale;
        double *orow = blk.out + (int64_t)(row0 + a) * blk.n_ket;
        for (int c = l; c < blk.n_ket; c += 32) {
          const double val = scale * entry_value(x, smax, mt->colrow, c0, blk.ket_masks[c]);
          orow[c] = val;
        }
      }
      WSYNC();
    }
  }
}

static size_t minors_smem_bytes(int nmax, int smax) {
  return sizeof(double) * ((size_t)smax * smax + (size_t)MB_WARPS * nmax * smax) + sizeof(WarpMeta) * MB_WARPS + 64;
}

}  // namespace tmf

extern "C" int64_t tmf_minor_desc_bytes(int nblocks) {
  return tmf::align256(64 * (int64_t)nblocks) + tmf::align256(4 * (int64_t)(nblocks + 1)) + 256;
}

extern "C" int tmf_minors_blocks(const tmf_minor_block *blocks_host, int nblocks, void *desc_dev,
                                 void *stream) {
  using namespace tmf;
  if (nblocks <= 0) return TMF_OK;
  std::vector<int> prefix(nblocks + 1, 0);
  int nmax = 1, smax = 1;
  for (int b = 0; b < nblocks; ++b) {
    const tmf_minor_block &k = blocks_host[b];
    if (k.s_bra > 64 || k.s_ket > 64 || k.minor > 32 || k.minor > k.s_ket || k.minor > k.s_bra) {
      set_error("tmf_minors_blocks: sometimes matrix > 64 or minor size > 32 not supported");
      return TMF_ERR_VALUE;
    }
    int ctas = (k.n_bra > 0 && k.n_ket > 0) ? (k.n_bra + MB_ROWS - 1) / MB_ROWS : 0;
    prefix[b + 1] = prefix[b] + ctas;
    nmax = std::max(nmax, k.minor);
    smax = std::max(smax, std::max(k.s_bra, k.s_ket));
  }
  if (prefix[nblocks] == 0) return TMF_OK;
  unsigned char *d = static_cast<unsigned char *>(desc_dev);
  const size_t o_pref = align256(sizeof(tmf_minor_block) * (size_t)nblocks);
  int rc = copy_h2d(d, blocks_host, sizeof(tmf_minor_block) * (size_t)nblocks, stream);
  if (rc) return rc;
  rc = copy_h2d(d + o_pref, prefix.data(), sizeof(int) * (size_t)(nblocks + 1), stream);
  if (rc) return rc;
  return launch_t("minors", minors_kernel, prefix[nblocks], 32 * tmf::MB_WARPS, minors_smem_bytes(nmax, smax), stream,
                reinterpret_cast<const tmf_minor_block *>(d), reinterpret_cast<const int *>(d + o_pref),
                nblocks, nmax, smax);
}
