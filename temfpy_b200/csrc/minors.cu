// K10: all minors det(S[rows(alpha)][:, cols(beta)]) of every charge block of every site.
//
// reference: slater.py:828-869 (_tensor_block: gather an (nsb, nsk, n, n) array and call a batched
// LAPACK det, one LU of size n ~ 11 per tensor entry; 87.5 % of the reference's run time) and the
// det_always scaling of :1137.
//
// Instead of one n x n LU per entry we share the elimination between all kets of a bra row:
//   X = S[rows(alpha), :]   (n x s_ket)  is row-reduced once (Gauss-Jordan, complete pivoting)
//   to [I | Y] on a pivot column set C0(alpha);  for any column set C with |C| = n
//      det X[:, C] = (prod of pivots) * sign * det Y[C0 \ C, C \ C0],
//   a determinant of size d = |C \ C0| (d <= 5 at chi = 1024, mean 2.5, versus n = 11).
// Complete pivoting keeps |Y| <= 1, so the small determinants are perfectly conditioned.
//
// One CTA handles up to MB_ROWS bra rows of one block: the row reductions of SLOTS rows proceed in
// lock-step through shared memory, then every thread evaluates one (alpha, beta) entry with the
// d x d matrix in registers and stores it coalesced along beta.
#include "cta.hpp"

namespace tmf {

constexpr int SLOTS = 8;       // bra rows reduced concurrently by one CTA
constexpr int MB_ROWS = 32;    // bra rows per CTA
constexpr int DMAX = 6;        // register path for reduced determinants up to 6 x 6
constexpr int DGEN = 16;       // generic local-memory path up to 16 x 16
static_assert(sizeof(tmf_minor_block) == 64, "block descriptor must be 64 bytes");

template <int D>
TMF_DEVICE double det_small(double (&m)[DMAX][DMAX]) {
  double det = 1.0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    int p = j;
    double best = fabs(m[j][j]);
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      double a = fabs(m[i][j]);
      if (a > best) { best = a; p = i; }
    }
    if (best == 0.0) return 0.0;
    if (p != j) {
      det = -det;
#pragma unroll
      for (int i = j + 1; i < D; ++i) {
        if (i == p) {
#pragma unroll
          for (int c = j; c < D; ++c) { double t = m[j][c]; m[j][c] = m[i][c]; m[i][c] = t; }
        }
      }
    }
    const double piv = m[j][j];
    det *= piv;
    const double inv = 1.0 / piv;
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      const double l = m[i][j] * inv;
#pragma unroll
      for (int c = j + 1; c < D; ++c) m[i][c] -= l * m[j][c];
    }
  }
  return det;
}

// generic fallback (d > DMAX): in-place LU on a local array
TMF_DEVICE double det_generic(double *m, int d, int ld) {
  double det = 1.0;
  for (int j = 0; j < d; ++j) {
    int p = j;
    double best = fabs(m[j * ld + j]);
    for (int i = j + 1; i < d; ++i)
      if (fabs(m[i * ld + j]) > best) { best = fabs(m[i * ld + j]); p = i; }
    if (best == 0.0) return 0.0;
    if (p != j) {
      det = -det;
      for (int c = 0; c < d; ++c) { double t = m[j * ld + c]; m[j * ld + c] = m[p * ld + c]; m[p * ld + c] = t; }
    }
    double piv = m[j * ld + j];
    det *= piv;
    for (int i = j + 1; i < d; ++i) {
      double l = m[i * ld + j] / piv;
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

// smem per slot: X (nmax x smax), plus bookkeeping
struct SlotMeta {
  uint64_t c0;        // pivot column set
  double scale;       // prod of pivots * sigma0 * det_always
  int colrow[64];     // pivot row of every pivot column
};

TMF_GLOBAL minors_kernel(const tmf_minor_block *blocks, const int *cta_prefix, int nblocks,
                         int nmax, int smax) {
  // locate (block, first row) of this CTA
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
  double *Ssm = reinterpret_cast<double *>(raw);             // sb * sk  (smax * smax)
  double *X = Ssm + (size_t)smax * smax;                       // SLOTS * nmax * smax
  double *red = X + (size_t)SLOTS * nmax * smax;               // SLOTS * 33
  SlotMeta *meta = reinterpret_cast<SlotMeta *>(red + SLOTS * 33);
  int *ired = reinterpret_cast<int *>(meta + SLOTS);           // SLOTS * 33 * 2
  int *rowdone = ired + SLOTS * 33 * 2;                        // SLOTS * 64 (pivot step of row, -1)
  int *pivrc = rowdone + SLOTS * 64;                           // SLOTS * 2

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

  for (int g0 = 0; g0 < nrows; g0 += SLOTS) {
    const int ns = (nrows - g0 < SLOTS) ? (nrows - g0) : SLOTS;
    // ---- gather X = S[rows(alpha), :] for the ns slots -----------------------------------
    PAR_FOR(idx, ns * sk) {
      int s = idx / sk, c = idx - s * sk;
      uint64_t rm = blk.bra_masks[row0 + g0 + s];
      double *x = X + (size_t)s * nmax * smax;
      int r = 0;
      while (rm) {
        int b = ctz64(rm);
        rm &= rm - 1;
        x[r * smax + c] = Ssm[(size_t)c * sb + b];
        ++r;
      }
    }
    PAR_FOR(idx, ns * 64) rowdone[idx] = -1;
    PAR_FOR(s, ns) {
      meta[s].c0 = 0;
      meta[s].scale = det_always;
    }
    CTA_SYNC();
    // ---- Gauss-Jordan with complete pivoting, n steps in lock-step -------------------------
    for (int t = 0; t < n; ++t) {
      PAR_FOR(item, ns * 32) {
        int s = item >> 5, lane = item & 31;
        const double *x = X + (size_t)s * nmax * smax;
        const uint64_t c0 = meta[s].c0;
        double best = -1.0;
        int br = 0, bc = 0;
        for (int e = lane; e < n * sk; e += 32) {
          int r = e / sk, c = e - r * sk;
          if (rowdone[s * 64 + r] >= 0 || ((c0 >> c) & 1)) continue;
          double a = fabs(x[r * smax + c]);
          if (a > best) { best = a; br = r; bc = c; }
        }
        red[s * 33 + lane] = best;
        ired[(s * 33 + lane) * 2] = br;
        ired[(s * 33 + lane) * 2 + 1] = bc;
      }
      CTA_SYNC();
      PAR_FOR(s, ns) {
        double best = red[s * 33];
        int br = ired[(s * 33) * 2], bc = ired[(s * 33) * 2 + 1];
        for (int l = 1; l < 32; ++l) {
          double v = red[s * 33 + l];
          int r = ired[(s * 33 + l) * 2], c = ired[(s * 33 + l) * 2 + 1];
          if (v > best || (v == best && (r < br || (r == br && c < bc)))) { best = v; br = r; bc = c; }
        }
        pivrc[s * 2] = br;
        pivrc[s * 2 + 1] = bc;
        double *x = X + (size_t)s * nmax * smax;
        double pv = (best >= 0.0) ? x[br * smax + bc] : 0.0;
        meta[s].scale *= pv;
        meta[s].c0 |= (1ull << bc);
        meta[s].colrow[bc] = br;
        rowdone[s * 64 + br] = t;
        red[s * 33 + 32] = (pv != 0.0) ? 1.0 / pv : 0.0;
      }
      CTA_SYNC();
      // normalise the pivot row
      PAR_FOR(idx, ns * sk) {
        int s = idx / sk, c = idx - s * sk;
        double *x = X + (size_t)s * nmax * smax;
        x[pivrc[s * 2] * smax + c] *= red[s * 33 + 32];
      }
      CTA_SYNC();
      // eliminate the pivot column from every other row (column-parallel: one thread per column)
      PAR_FOR(idx, ns * sk) {
        int s = idx / sk, c = idx - s * sk;
        double *x = X + (size_t)s * nmax * smax;
        const int pr = pivrc[s * 2], pc = pivrc[s * 2 + 1];
        if (c != pc) {
          const double u = x[pr * smax + c];
          for (int r = 0; r < n; ++r)
            if (r != pr) x[r * smax + c] -= x[r * smax + pc] * u;
        }
      }
      CTA_SYNC();
      PAR_FOR(idx, ns * n) {  // the pivot column itself becomes a unit vector
        int s = idx / n, r = idx - s * n;
        double *x = X + (size_t)s * nmax * smax;
        const int pr = pivrc[s * 2], pc = pivrc[s * 2 + 1];
        x[r * smax + pc] = (r == pr) ? 1.0 : 0.0;
      }
      CTA_SYNC();
    }
    // sigma0: sign of the permutation (rank of pivot column) -> pivot row
    PAR_FOR(s, ns) {
      uint64_t c0 = meta[s].c0;
      uint64_t seen = 0;
      int inv = 0;
      while (c0) {
        int c = ctz64(c0);
        c0 &= c0 - 1;
        int r = meta[s].colrow[c];
        inv += popc64(seen >> (r + 1));   // earlier columns mapped to larger rows
        seen |= (1ull << r);
      }
      if (inv & 1) meta[s].scale = -meta[s].scale;
    }
    CTA_SYNC();
    // ---- one thread per (alpha, beta) entry -------------------------------------------------
    PAR_FOR(idx, ns * blk.n_ket) {
      const int s = idx / blk.n_ket, c = idx - s * blk.n_ket;
      const double *x = X + (size_t)s * nmax * smax;
      const uint64_t c0 = meta[s].c0;
      const uint64_t cm = blk.ket_masks[c];
      const uint64_t U = cm & c0;
      uint64_t mu = c0 & ~cm, de = cm & ~c0;
      const int d = popc64(de);
      int par = 0;
      double val;
      if (d <= DMAX) {
        int rr[DMAX], cc[DMAX];
#pragma unroll
        for (int i = 0; i < DMAX; ++i) {
          rr[i] = 0;
          cc[i] = 0;
          if (i < d) {
            int m = ctz64(mu), e = ctz64(de);
            mu &= mu - 1;
            de &= de - 1;
            int a = m < e ? m : e, b = m < e ? e : m;
            uint64_t between = (b - a > 1) ? (((1ull << (b - a - 1)) - 1) << (a + 1)) : 0ull;
            par += popc64(U & between);
            rr[i] = meta[s].colrow[m];
            cc[i] = e;
          }
        }
        double m8[DMAX][DMAX];
#pragma unroll
        for (int i = 0; i < DMAX; ++i)
#pragma unroll
          for (int j = 0; j < DMAX; ++j)
            m8[i][j] = (i < d && j < d) ? x[rr[i] * smax + cc[j]] : ((i == j) ? 1.0 : 0.0);
        switch (d) {
          case 0: val = 1.0; break;
          case 1: val = m8[0][0]; break;
          case 2: val = m8[0][0] * m8[1][1] - m8[0][1] * m8[1][0]; break;
          case 3: val = det_small<3>(m8); break;
          case 4: val = det_small<4>(m8); break;
          case 5: val = det_small<5>(m8); break;
          default: val = det_small<6>(m8); break;
        }
      } else if (d <= DGEN) {
        double buf[DGEN * DGEN];
        int rows_[DGEN], cols_[DGEN];
        for (int i = 0; i < d; ++i) {
          int m = ctz64(mu), e = ctz64(de);
          mu &= mu - 1;
          de &= de - 1;
          int a = m < e ? m : e, b = m < e ? e : m;
          uint64_t between = (b - a > 1) ? (((1ull << (b - a - 1)) - 1) << (a + 1)) : 0ull;
          par += popc64(U & between);
          rows_[i] = meta[s].colrow[m];
          cols_[i] = e;
        }
        for (int i = 0; i < d; ++i)
          for (int j = 0; j < d; ++j) buf[i * DGEN + j] = x[rows_[i] * smax + cols_[j]];
        val = det_generic(buf, d, DGEN);
      } else {
        val = NAN;  // would need > 16 simultaneous column exchanges; not produced by Schmidt sets
      }
      val *= meta[s].scale;
      if (par & 1) val = -val;
      blk.out[(int64_t)(row0 + g0 + s) * blk.n_ket + c] = val;
    }
    CTA_SYNC();
  }
}

static size_t minors_smem_bytes(int nmax, int smax) {
  return sizeof(double) * ((size_t)smax * smax + (size_t)SLOTS * nmax * smax + SLOTS * 33) +
         sizeof(SlotMeta) * SLOTS + sizeof(int) * (SLOTS * 33 * 2 + SLOTS * 64 + SLOTS * 2 + 8);
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
  return launch(minors_kernel, prefix[nblocks], 256, minors_smem_bytes(nmax, smax), stream,
                reinterpret_cast<const tmf_minor_block *>(d), reinterpret_cast<const int *>(d + o_pref),
                nblocks, nmax, smax);
}
