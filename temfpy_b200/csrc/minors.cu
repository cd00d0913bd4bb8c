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

// Exclusive prefix parity of a mask: bit j = parity of the number of set bits below j.  With it the
// sign of the column exchange C0 -> C (the permutation that sorts U + E after the columns M of C0 were
// replaced in place by E) is  popc(M & pp(C0)) + popc(E & pp(C))  mod 2  -- two ANDs and two POPCs per
// entry instead of one masked popcount per exchanged column.
template <typename MT>
TMF_DEVICE MT prefix_parity(MT x) {
  MT p = (MT)(x << 1);
  p ^= (MT)(p << 1);
  p ^= (MT)(p << 2);
  p ^= (MT)(p << 4);
  p ^= (MT)(p << 8);
  p ^= (MT)(p << 16);
  if (sizeof(MT) == 8) p ^= (MT)((uint64_t)p << 32);
  return p;
}
TMF_DEVICE int popcm(uint32_t x) {
#if defined(TMF_HOSTSIM)
  return __builtin_popcount(x);
#else
  return __popc(x);
#endif
}
TMF_DEVICE int popcm(uint64_t x) { return popc64(x); }
TMF_DEVICE int ctzm(uint32_t x) {
#if defined(TMF_HOSTSIM)
  return __builtin_ctz(x);
#else
  return __ffs((int)x) - 1;
#endif
}
TMF_DEVICE int ctzm(uint64_t x) { return ctz64(x); }

// One tensor entry with exactly D exchanged columns (D = 1..4): gathers Y[C0 \ C, C \ C0] (D x D) from
// the reduced matrix and returns the signed determinant.  D is a compile-time constant: the entries of a
// bra row are binned by D first (see the kernel), so every lane of a warp runs the same straight-line code.
template <int D, typename MT>
TMF_DEVICE double entry_d(const double *x, int smax, const int *colrow, MT c0, MT pp0, MT cm, MT ppk) {
  MT mu = c0 & ~cm, de = cm & ~c0;
  const int par = popcm((MT)(mu & pp0)) ^ popcm((MT)(de & ppk));
  int rr[D], cc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int m = ctzm(mu), e = ctzm(de);
    mu &= (MT)(mu - 1);
    de &= (MT)(de - 1);
    rr[i] = colrow[m] * smax;
    cc[i] = e;
  }
  double val;
  if constexpr (D == 1) {
    val = x[rr[0] + cc[0]];
  } else if constexpr (D == 2) {
    val = det2(x[rr[0] + cc[0]], x[rr[0] + cc[1]], x[rr[1] + cc[0]], x[rr[1] + cc[1]]);
  } else if constexpr (D == 3) {
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
  return (par & 1) ? -val : val;
}

// any number of exchanged columns (dispatch on d; used by the simulator path and for d > 4)
template <typename MT>
TMF_DEVICE double entry_any(const double *x, int smax, const int *colrow, MT c0, MT pp0, MT cm, MT ppk) {
  const MT de = cm & ~c0;
  const int d = popcm(de);
  switch (d) {
    case 0: return 1.0;
    case 1: return entry_d<1, MT>(x, smax, colrow, c0, pp0, cm, ppk);
    case 2: return entry_d<2, MT>(x, smax, colrow, c0, pp0, cm, ppk);
    case 3: return entry_d<3, MT>(x, smax, colrow, c0, pp0, cm, ppk);
    case 4: return entry_d<4, MT>(x, smax, colrow, c0, pp0, cm, ppk);
    default: break;
  }
  int par = 0;
  const double val = entry_generic(x, smax, colrow, (uint64_t)(cm & c0), (uint64_t)(c0 & ~cm), (uint64_t)de, d, par);
  return (par & 1) ? -val : val;
}

// ---------------------------------------------------------------------------------------------
// Warp-independent kernel: after the sometimes matrix and the ket masks are staged in shared memory (one
// CTA barrier) every warp processes bra rows on its own -- row reduction with lanes = columns in a
// private shared-memory tile, then the entries of the row, binned by their number of exchanged columns
// so that each bin runs divergence-free -- synchronising with __syncwarp only.
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
  uint64_t pp0;       // its exclusive prefix parity
  double scale;       // prod of pivots * sigma0 * det_always
  double inv;         // 1 / current pivot
  int pc, pad_;       // current pivot column
  int colrow[64];     // pivot row of every pivot column
  double cand[64];    // simulator only: candidates of the pivot search
};

constexpr int MB_WARPS = 8;
constexpr int NCLS = 5;        // bins: 1, 2, 3, 4 exchanged columns, more
constexpr int QCAP = 64;       // a bin is flushed as soon as it holds a full warp of entries

#if !defined(TMF_HOSTSIM)
// processes the entries q[i0 .. i0 + cnt) (cnt <= 32) of one bin
template <int D, typename MT>
__device__ __forceinline__ void flush_bin(const unsigned short *q, int i0, int cnt, int lane, const double *x, int smax,
                                          const int *colrow, MT c0, MT pp0, const MT *kmask, const MT *kpp,
                                          double scale, double *orow) {
  if (lane < cnt) {
    const int c = q[i0 + lane];
    double v;
    if constexpr (D <= 4) v = entry_d<D, MT>(x, smax, colrow, c0, pp0, kmask[c], kpp[c]);
    else v = entry_any<MT>(x, smax, colrow, c0, pp0, kmask[c], kpp[c]);
    orow[c] = scale * v;
  }
}
#endif

#if !defined(TMF_HOSTSIM)
// Register-resident Gauss-Jordan of one bra row for sometimes matrices of up to 32 columns and minors
// of up to NR rows: lane c keeps column c of X in registers, the pivot column is broadcast through NR
// doubles of shared memory, the pivot search is one REDUX.  Every lane tracks the pivot set, the
// permutation parity and the pivot product redundantly (uniform registers), so no lane-0 section and no
// shared-memory round trip of the matrix remain; the reduced matrix is written to the warp's tile once,
// for the entry gathers.  Same pivoting rule and arithmetic as the shared-memory path below.
template <int NR>
__device__ __forceinline__ void reduce_row_regs(const double *Ssm, int ldS, int sk, int n, uint64_t rmask,
                                                int lane, double *x, int smax, int *colrow, double *bcast,
                                                double det_always, uint32_t &c0_out, double &scale_out) {
  double xr[NR];
  {
    uint64_t rm = rmask;
    const double *col = Ssm + (size_t)(lane < sk ? lane : 0) * ldS;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double v = 0.0;
      if (r < n) {
        const int bit = ctz64(rm);
        rm &= rm - 1;
        if (lane < sk) v = col[bit];
      }
      xr[r] = v;
    }
  }
  uint32_t c0 = 0u;
  double scale = det_always;
  int invs = 0;
#pragma unroll
  for (int t = 0; t < NR; ++t) {
    if (t < n) {
      const double v = xr[t];
      // largest |v| among the unused columns: float bit pattern (monotone), bit 5 marks a candidate,
      // bits 0-4 carry 31 - column so that ties go to the lowest column
      unsigned key = 0u;
      if (lane < sk && !((c0 >> lane) & 1u))
        key = (__float_as_uint(fabsf((float)v)) & ~63u) | 32u | (unsigned)(31 - lane);
      key = __reduce_max_sync(0xffffffffu, key);
      const int pc = 31 - (int)(key & 31u);
      const double pv = __shfl_sync(0xffffffffu, v, pc);
      const double inv = (pv != 0.0) ? 1.0 / pv : 0.0;
      scale *= pv;
      invs += __popc((unsigned)((uint64_t)c0 >> (pc + 1)));
      c0 |= 1u << pc;
      if (lane == pc) {
        colrow[pc] = t;
#pragma unroll
        for (int r = 0; r < NR; r += 2) *reinterpret_cast<double2 *>(bcast + r) = make_double2(xr[r], xr[r + 1]);
      }
      __syncwarp();
      const double u = (lane == pc) ? 0.0 : v * inv;
#pragma unroll
      for (int r = 0; r < NR; r += 2) {
        const double2 b = *reinterpret_cast<const double2 *>(bcast + r);
        if (r != t) xr[r] -= b.x * u;
        if (r + 1 != t) xr[r + 1] -= b.y * u;
      }
      xr[t] = u;
      __syncwarp();
    }
  }
  if (lane < sk) {
#pragma unroll
    for (int r = 0; r < NR; ++r)
      if (r < n) x[r * smax + lane] = xr[r];
  }
  c0_out = c0;
  scale_out = (invs & 1) ? -scale : scale;
}
#endif

// STAGE: the entries of a bra row are collected in a shared-memory row buffer of the warp and written out as
// whole 256-byte warp stores.  Used when `out` is a peer window of another GPU (multi-GPU gather fused into this
// kernel, dist.FusedGather): the bins below produce the entries of a row in scattered order, which the local L2
// merges but NVLink would carry as 8-byte packets.
template <typename MT, bool STAGE>
TMF_GLOBAL_LB(256, 3) minors_kernel(const tmf_minor_block *blocks, const int *cta_prefix, int nblocks,
                         int nmax, int smax, int nkmax) {
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
  const int ldS = sb | 1;                                        // odd leading dimension: conflict-free column gathers
  double *Ssm = reinterpret_cast<double *>(raw);                 // smax * (smax | 1)
  double *Xall = Ssm + (size_t)smax * (smax | 1);                // MB_WARPS * nmax * smax
  double *bcast_all = Ssm + ((((size_t)smax * (smax | 1) + (size_t)MB_WARPS * nmax * smax) + 1) & ~(size_t)1);   // MB_WARPS * 16, 16-byte aligned
  WarpMeta *metas = reinterpret_cast<WarpMeta *>(bcast_all + MB_WARPS * 16);
  MT *kmask = reinterpret_cast<MT *>(metas + MB_WARPS);          // nkmax
  MT *kpp = kmask + nkmax;                                       // nkmax
  unsigned short *queues = reinterpret_cast<unsigned short *>(kpp + nkmax);   // MB_WARPS * NCLS * QCAP
  double *rows_all = reinterpret_cast<double *>(
      (reinterpret_cast<uintptr_t>(queues + (size_t)MB_WARPS * NCLS * QCAP) + 7) & ~uintptr_t(7));   // STAGE: MB_WARPS * nkmax

  const double det_always = blk.det ? *blk.det : 1.0;
  PAR_FOR(idx, sb * sk) {
    const int c = idx / sb, r = idx - c * sb;
    Ssm[c * ldS + r] = blk.S[idx];
  }
  PAR_FOR(c, blk.n_ket) {
    const MT km = (MT)blk.ket_masks[c];
    kmask[c] = km;
    kpp[c] = prefix_parity<MT>(km);
  }
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
#if !defined(TMF_HOSTSIM)
      if (sizeof(MT) == 4 && n <= 16) {   // (uniform) register-resident reduction
        const int lane = threadIdx.x & 31;
        uint32_t c0r;
        double sc;
        double *bc = bcast_all + w * 16;
        if (n <= 8) reduce_row_regs<8>(Ssm, ldS, sk, n, rmask, lane, x, smax, mt->colrow, bc, det_always, c0r, sc);
        else if (n <= 12) reduce_row_regs<12>(Ssm, ldS, sk, n, rmask, lane, x, smax, mt->colrow, bc, det_always, c0r, sc);
        else reduce_row_regs<16>(Ssm, ldS, sk, n, rmask, lane, x, smax, mt->colrow, bc, det_always, c0r, sc);
        if (lane == 0) {
          mt->c0 = c0r;
          mt->pp0 = prefix_parity<uint32_t>(c0r);
          mt->scale = sc;
        }
        __syncwarp();
      } else {
#endif
      // ---- gather X = S[rows(alpha), :] (lane = column) -------------------------------------
      LANE_FOR(l) {
        for (int c = l; c < sk; c += 32) {
          uint64_t rm = rmask;
          int r = 0;
          while (rm) {
            const int bit = ctz64(rm);
            rm &= rm - 1;
            x[r * smax + c] = Ssm[(size_t)c * ldS + bit];
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
      // sigma0: sign of the permutation (rank of pivot column) -> pivot row; prefix parity of C0
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
        mt->pp0 = prefix_parity<uint64_t>(mt->c0);
      }
      WSYNC();
#if !defined(TMF_HOSTSIM)
      }
#endif
      // ---- the entries of the row -------------------------------------------------------------
      double *grow = blk.out + (int64_t)(row0 + a) * blk.n_ket;
      double *orow = STAGE ? rows_all + (size_t)w * nkmax : grow;
#if defined(TMF_HOSTSIM)
      orow = grow;
      LANE_FOR(l) {
        const MT c0 = (MT)mt->c0, pp0 = (MT)mt->pp0;
        for (int c = l; c < blk.n_ket; c += 32)
          orow[c] = mt->scale * entry_any<MT>(x, smax, mt->colrow, c0, pp0, kmask[c], kpp[c]);
      }
#else
      {
        const int lane = threadIdx.x & 31;
        const unsigned lt = (1u << lane) - 1u;
        const MT c0 = (MT)mt->c0, pp0 = (MT)mt->pp0;
        const double scale = mt->scale;
        const int *colrow = mt->colrow;
        unsigned short *q = queues + (size_t)w * NCLS * QCAP;
        int cnt[NCLS];
#pragma unroll
        for (int k = 0; k < NCLS; ++k) cnt[k] = 0;
        for (int cb = 0; cb < blk.n_ket; cb += 32) {
          const int c = cb + lane;
          int cls = 0;
          if (c < blk.n_ket) {
            const int d = popcm((MT)(kmask[c] & ~c0));
            if (d == 0) orow[c] = scale;
            cls = d < NCLS ? d : NCLS;
          }
#pragma unroll
          for (int k = 0; k < NCLS; ++k) {
            const unsigned bal = __ballot_sync(0xffffffffu, cls == k + 1);
            if (cls == k + 1) q[k * QCAP + cnt[k] + __popc(bal & lt)] = (unsigned short)c;
            cnt[k] += __popc(bal);
          }
          __syncwarp();
          if (cnt[0] >= 32) { cnt[0] -= 32; flush_bin<1, MT>(q + 0 * QCAP, cnt[0], 32, lane, x, smax, colrow, c0, pp0, kmask, kpp, scale, orow); }
          if (cnt[1] >= 32) { cnt[1] -= 32; flush_bin<2, MT>(q + 1 * QCAP, cnt[1], 32, lane, x, smax, colrow, c0, pp0, kmask, kpp, scale, orow); }
          if (cnt[2] >= 32) { cnt[2] -= 32; flush_bin<3, MT>(q + 2 * QCAP, cnt[2], 32, lane, x, smax, colrow, c0, pp0, kmask, kpp, scale, orow); }
          if (cnt[3] >= 32) { cnt[3] -= 32; flush_bin<4, MT>(q + 3 * QCAP, cnt[3], 32, lane, x, smax, colrow, c0, pp0, kmask, kpp, scale, orow); }
          if (cnt[4] >= 32) { cnt[4] -= 32; flush_bin<5, MT>(q + 4 * QCAP, cnt[4], 32, lane, x, smax, colrow, c0, pp0, kmask, kpp, scale, orow); }
          __syncwarp();
        }
        if (cnt[0]) flush_bin<1, MT>(q + 0 * QCAP, 0, cnt[0], lane, x, smax, colrow, c0, pp0, kmask, kpp, scale, orow);
        if (cnt[1]) flush_bin<2, MT>(q + 1 * QCAP, 0, cnt[1], lane, x, smax, colrow, c0, pp0, kmask, kpp, scale, orow);
        if (cnt[2]) flush_bin<3, MT>(q + 2 * QCAP, 0, cnt[2], lane, x, smax, colrow, c0, pp0, kmask, kpp, scale, orow);
        if (cnt[3]) flush_bin<4, MT>(q + 3 * QCAP, 0, cnt[3], lane, x, smax, colrow, c0, pp0, kmask, kpp, scale, orow);
        if (cnt[4]) flush_bin<5, MT>(q + 4 * QCAP, 0, cnt[4], lane, x, smax, colrow, c0, pp0, kmask, kpp, scale, orow);
        if (STAGE) {
          __syncwarp();
          for (int c = lane; c < blk.n_ket; c += 32) grow[c] = orow[c];
        }
      }
#endif
      WSYNC();
    }
  }
}

static size_t minors_smem_bytes(int nmax, int smax, int nkmax, size_t mask_bytes, bool stage) {
  return sizeof(double) * ((size_t)smax * (smax | 1) + (size_t)MB_WARPS * nmax * smax + MB_WARPS * 16 + 2) +
         sizeof(WarpMeta) * MB_WARPS +
         2 * mask_bytes * (size_t)nkmax + sizeof(unsigned short) * MB_WARPS * NCLS * QCAP + 64 +
         (stage ? sizeof(double) * (size_t)MB_WARPS * nkmax + 8 : 0);
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
  int nmax = 1, smax = 1, nkmax = 1;
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
    nkmax = std::max(nkmax, k.n_ket);
  }
  nkmax = (nkmax + 3) & ~3;
  if (nkmax > 65535) {
    set_error("tmf_minors_blocks: more than 65535 Schmidt vectors in one charge sector");
    return TMF_ERR_VALUE;
  }
  if (prefix[nblocks] == 0) return TMF_OK;
  unsigned char *d = static_cast<unsigned char *>(desc_dev);
  const size_t o_pref = align256(sizeof(tmf_minor_block) * (size_t)nblocks);
  int rc = copy_h2d(d, blocks_host, sizeof(tmf_minor_block) * (size_t)nblocks, stream);
  if (rc) return rc;
  rc = copy_h2d(d + o_pref, prefix.data(), sizeof(int) * (size_t)(nblocks + 1), stream);
  if (rc) return rc;
  // occupation masks fit 32 bits for the usual sometimes matrices (<= 32 columns): half the integer work
  const bool stage = (blocks_host[0].pad_ & 1) != 0;      // out is a peer window: row-staged 256-byte stores
  const tmf_minor_block *bd = reinterpret_cast<const tmf_minor_block *>(d);
  const int *pd = reinterpret_cast<const int *>(d + o_pref);
  const int grid = prefix[nblocks], thr = 32 * tmf::MB_WARPS;
  if (smax <= 32) {
    const size_t sm = minors_smem_bytes(nmax, smax, nkmax, 4, stage);
    return stage ? launch_t("minors", minors_kernel<uint32_t, true>, grid, thr, sm, stream, bd, pd, nblocks, nmax, smax, nkmax)
                 : launch_t("minors", minors_kernel<uint32_t, false>, grid, thr, sm, stream, bd, pd, nblocks, nmax, smax, nkmax);
  }
  const size_t sm = minors_smem_bytes(nmax, smax, nkmax, 8, stage);
  return stage ? launch_t("minors", minors_kernel<uint64_t, true>, grid, thr, sm, stream, bd, pd, nblocks, nmax, smax, nkmax)
               : launch_t("minors", minors_kernel<uint64_t, false>, grid, thr, sm, stream, bd, pd, nblocks, nmax, smax, nkmax);
}
