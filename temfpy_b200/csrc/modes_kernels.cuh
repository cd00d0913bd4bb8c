// CTA kernels of the Schmidt-mode extraction (K3/K4): panel orthonormalisation, one-sided Jacobi
// (small SVD / symmetric eigenproblem in shared memory), pivoted Cholesky of the filled-space
// projector, and the direct small-block eigen-solver.  All written in the PAR_FOR/CTA_SYNC
// style of cta.hpp so that the CPU simulator executes the very same source.
#pragma once
#include "cta.hpp"
#include "jacobi.cuh"
#include <cstdio>
#include <cstdlib>

namespace tmf {

constexpr int PANEL_W = 16;        // panel width of the block Gram-Schmidt
constexpr int SMALL_N_MAX = 64;    // largest block the direct solver's shared-memory layout is sized for
// Blocks up to this size are diagonalised directly (one CTA, Jacobi: ~2 ms for n = 64, 0.5 ms for n = 32), larger
// ones go through the sketch path.  The nested chain driver only needs the entangled modes and uses 32; the
// legacy form (explicit filled bases; iMPS / Pfaffian drivers) keeps 64.  TMF_SMALL_N overrides both.
inline int small_n(bool nested) {
  static const int env = [] {
    const char *e = std::getenv("TMF_SMALL_N");
    const int x = e ? std::atoi(e) : 0;
    return x < 0 ? 0 : (x > SMALL_N_MAX ? SMALL_N_MAX : x);
  }();
  return env > 0 ? env : (nested ? 32 : 64);
}
constexpr int JAC_SMEM_J_MAX = 96; // above this the Jacobi rotation matrix lives in global memory
constexpr int PIVCHOL_MAX_PARTS = 8;
constexpr int R_SKETCH_MAX = 160;  // G (r x r) must fit in shared memory

// ---------------------------------------------------------------------------------------------
// column norms (squared) of a column-major matrix: one CTA per job
// ---------------------------------------------------------------------------------------------
struct NormJob {
  const double *Y;
  double *out;  // ncols
  int rows, ld, ncols, pad_;
};
TMF_GLOBAL colnorm_kernel(const NormJob *jobs) {
  const NormJob jb = jobs[BLOCK_ID];
  DYN_SMEM(double, part);  // ncols * 33
  PAR_FOR(item, jb.ncols * 32) {
    int c = item >> 5, lane = item & 31;
    const double *y = jb.Y + (int64_t)c * jb.ld;
    double s = 0.0;
    for (int r = lane; r < jb.rows; r += 32) s += y[r] * y[r];
    part[c * 33 + lane] = s;
  }
  CTA_SYNC();
  PAR_FOR(c, jb.ncols) {
    double s = 0.0;
    for (int l = 0; l < 32; ++l) s += part[c * 33 + l];
    jb.out[c] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// panel MGS2: orthonormalises ncols (<= PANEL_W) columns in place.  Columns whose norm falls
// below rel_tol * (norm before any projection) are numerically dependent and are set to zero.
// ---------------------------------------------------------------------------------------------
struct PanelJob {
  double *P;            // rows x ncols, column-major, ld
  const double *norm0;  // squared norms of the original columns (ncols)
  int *nzero;           // incremented for every zeroed column (may be null)
  const double *Qprev;  // previous (orthonormal) panels, rows x c0, same ld (fused projection; may be null)
  int rows, ld, ncols, use_smem;
  int c0, pad_;         // number of previous columns to project out before the panel step (0: none)
};

TMF_DEVICE void panel_mgs2_body(const PanelJob &jb, double rel_tol2, double *sm) {
  double *part = sm;                        // PANEL_W * 33
  double *coef = part + PANEL_W * 33;       // PANEL_W
  double *red = coef + PANEL_W;             // 32 + 2
  double *panel = red + 40;                 // rows * ncols when use_smem
  double *P = jb.P;
  int ld = jb.ld;
  if (jb.use_smem) {
    PAR_FOR(idx, jb.rows * jb.ncols) {
      int c = idx / jb.rows, r = idx - c * jb.rows;
      panel[idx] = jb.P[(int64_t)c * jb.ld + r];
    }
    CTA_SYNC();
    P = panel;
    ld = jb.rows;
  }
  for (int j = 0; j < jb.ncols; ++j) {
    double *pj = P + (int64_t)j * ld;
    for (int pass = 0; pass < 2 && j > 0; ++pass) {
      PAR_FOR(item, j * 32) {
        int i = item >> 5, lane = item & 31;
        const double *pi = P + (int64_t)i * ld;
        double s = 0.0;
        for (int r = lane; r < jb.rows; r += 32) s += pi[r] * pj[r];
        part[i * 33 + lane] = s;
      }
      CTA_SYNC();
      PAR_FOR(i, j) {
        double s = 0.0;
        for (int l = 0; l < 32; ++l) s += part[i * 33 + l];
        coef[i] = s;
      }
      CTA_SYNC();
      PAR_FOR(r, jb.rows) {
        double v = pj[r];
        for (int i = 0; i < j; ++i) v -= coef[i] * P[(int64_t)i * ld + r];
        pj[r] = v;
      }
      CTA_SYNC();
    }
    PAR_FOR(lane, 32) {
      double s = 0.0;
      for (int r = lane; r < jb.rows; r += 32) s += pj[r] * pj[r];
      red[lane] = s;
    }
    CTA_SYNC();
    PAR_FOR(one, 1) {
      double s = 0.0;
      for (int l = 0; l < 32; ++l) s += red[l];
      double scale = 0.0;
      if (s > rel_tol2 * jb.norm0[j] && s > 0.0) scale = 1.0 / sqrt(s);
      else if (jb.nzero) *jb.nzero += 1;
      red[32] = scale;
    }
    CTA_SYNC();
    const double scale = red[32];
    PAR_FOR(r, jb.rows) pj[r] *= scale;
    CTA_SYNC();
  }
  if (jb.use_smem) {
    PAR_FOR(idx, jb.rows * jb.ncols) {
      int c = idx / jb.rows, r = idx - c * jb.rows;
      jb.P[(int64_t)c * jb.ld + r] = panel[idx];
    }
  }
}
TMF_GLOBAL panel_mgs2_kernel(const PanelJob *jobs, double rel_tol2) {
  const PanelJob jb = jobs[BLOCK_ID];
  if (jb.rows <= 0 || jb.ncols <= 0) return;
  DYN_SMEM(double, sm);
  panel_mgs2_body(jb, rel_tol2, sm);
}

// ---------------------------------------------------------------------------------------------
// panel Cholesky-QR: G = P^T P (w x w), G = R^T R, P <- P R^-1.  Everything is parallel over the rows
// (the MGS2 kernel above walks the columns one by one: ~100 dependent CTA phases per panel), so a panel
// costs a few microseconds.  A sketch panel after the projection on the previous panels has a condition
// number ~1e5 (16 columns of a geometrically decaying spectrum): one pass leaves ~kappa^2 eps of
// non-orthogonality, which the second round of the block Gram-Schmidt driver removes (CholQR2).  If a pivot
// of the Cholesky factorisation collapses (kappa > ~1e6, dependent or zero columns) the CTA falls back to
// the MGS2 body on the panel in global memory.
// ---------------------------------------------------------------------------------------------
#if defined(TMF_HOSTSIM)
static const double CHOLQR_PIVOT_TOL = std::getenv("TMF_CHOLQR_TOL") ? std::atof(std::getenv("TMF_CHOLQR_TOL")) : 1e-12;
#else
constexpr double CHOLQR_PIVOT_TOL = 1e-12;
#endif
TMF_GLOBAL_LB(256, 3) panel_cholqr_kernel(const PanelJob *jobs, double rel_tol2) {
  const PanelJob jb0 = jobs[BLOCK_ID];
  if (jb0.rows <= 0 || jb0.ncols <= 0) return;
  const int w = jb0.ncols, rows = jb0.rows, ld = jb0.ld;
  double *P = jb0.P;
  DYN_SMEM(double, sm);
  double *G = sm;                      // PANEL_W * PANEL_W
  double *Ri = G + PANEL_W * PANEL_W;  // inverse of R (upper triangular), Ri[k * PANEL_W + j]
  int *flag = reinterpret_cast<int *>(Ri + PANEL_W * PANEL_W);
  double *scratch = Ri + PANEL_W * PANEL_W + 2;   // MGS2 fallback scratch
  // ---- fused block Gram-Schmidt projection: P -= Qprev (Qprev^T P) ---------------------------------
  // (two skinny GEMMs per panel and round in the previous version: 64 x 64 DMMA tiles of which a 16- or
  // 32-row by 16-column corner was useful, and two more launches in the dependent chain)
  const int c0 = (jb0.Qprev != nullptr) ? jb0.c0 : 0;
  if (c0 > 0) {
    double *coef = scratch + PANEL_W * 33 + PANEL_W + 40 + 8;   // c0 x PANEL_W, coef[i * PANEL_W + j]
    const double *Q = jb0.Qprev;
#if !defined(TMF_HOSTSIM)
    {
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
      const int nbi = (c0 + 3) / 4, nbj = (w + 3) / 4;
      for (int blk = wid; blk < nbi * nbj; blk += nw) {
        const int bi = blk / nbj, bj = blk - bi * nbj;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        for (int r = lane; r < rows; r += 32) {
          double qi[4], pj[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            qi[a] = (4 * bi + a < c0) ? Q[(int64_t)(4 * bi + a) * ld + r] : 0.0;
            pj[a] = (4 * bj + a < w) ? P[(int64_t)(4 * bj + a) * ld + r] : 0.0;
          }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] += qi[a] * pj[b];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            double v = acc[a][b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && 4 * bi + a < c0) coef[(4 * bi + a) * PANEL_W + 4 * bj + b] = (4 * bj + b < w) ? v : 0.0;
          }
      }
    }
#else
    PAR_FOR(idx, c0 * PANEL_W) {
      const int i = idx / PANEL_W, j = idx - i * PANEL_W;
      double sacc = 0.0;
      if (j < w)
        for (int r = 0; r < rows; ++r) sacc += Q[(int64_t)i * ld + r] * P[(int64_t)j * ld + r];
      coef[idx] = sacc;
    }
#endif
    CTA_SYNC();
    PAR_FOR(r, rows) {
      double p[PANEL_W];
#pragma unroll
      for (int j = 0; j < PANEL_W; ++j) p[j] = (j < w) ? P[(int64_t)j * ld + r] : 0.0;
      for (int i = 0; i < c0; ++i) {
        const double q = Q[(int64_t)i * ld + r];
        const double *ci = coef + i * PANEL_W;
#pragma unroll
        for (int j = 0; j < PANEL_W; ++j) p[j] -= q * ci[j];
      }
#pragma unroll
      for (int j = 0; j < PANEL_W; ++j)
        if (j < w) P[(int64_t)j * ld + r] = p[j];
    }
    CTA_SYNC();
  }
#if !defined(TMF_HOSTSIM)
  {
    // Gram matrix: warp -> 4 x 4 blocks of column pairs (register tile), lanes -> rows, shuffle reduction
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nbk = (w + 3) / 4;
    for (int blk = wid; blk < nbk * nbk; blk += nw) {
      const int bi = blk / nbk, bj = blk - bi * nbk;
      if (bj < bi) continue;             // symmetric: upper blocks only
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
      for (int r = lane; r < rows; r += 32) {
        double pi[4], pj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          pi[a] = (4 * bi + a < w) ? P[(int64_t)(4 * bi + a) * ld + r] : 0.0;
          pj[a] = (4 * bj + a < w) ? P[(int64_t)(4 * bj + a) * ld + r] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] += pi[a] * pj[b];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          double v = acc[a][b];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0 && 4 * bi + a < w && 4 * bj + b < w) {
            G[(4 * bi + a) * PANEL_W + 4 * bj + b] = v;
            G[(4 * bj + b) * PANEL_W + 4 * bi + a] = v;
          }
        }
    }
  }
#else
  PAR_FOR(idx, w * w) {
    const int i = idx / w, j = idx - i * w;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += P[(int64_t)i * ld + r] * P[(int64_t)j * ld + r];
    G[i * PANEL_W + j] = s;
  }
#endif
  CTA_SYNC();
#if !defined(TMF_HOSTSIM)
  // Cholesky G = L L^T (R = L^T) by warp 0: lane i keeps row i of the lower triangle in registers, the
  // pivot row is broadcast by shuffles (left-looking, one column per step); then lane j inverts column j.
  if (threadIdx.x < 32) {
    const int i = threadIdx.x;
    double g[PANEL_W];
#pragma unroll
    for (int q = 0; q < PANEL_W; ++q) g[q] = (i < w && q < w) ? G[i * PANEL_W + q] : ((i == q) ? 1.0 : 0.0);
    double diag0 = 0.0;
#pragma unroll
    for (int q = 0; q < PANEL_W; ++q) if (q == i) diag0 = g[q];
    // a column whose norm after the projection is below the rounding noise of the original sketch column
    // carries no direction of its own (the range of B is exhausted: normalising it would put a copy of
    // directions already in Q back into the basis).  Such columns are zeroed: their row / column of G
    // becomes that of the identity, so the factorisation passes over them.
    const double noise0 = (i < w) ? rel_tol2 * jb0.norm0[i] : 0.0;
    const unsigned zmask = __ballot_sync(0xffffffffu, i < w && !(diag0 > noise0));
    if (zmask) {
      const bool zi = (zmask >> i) & 1u;
#pragma unroll
      for (int q = 0; q < PANEL_W; ++q)
        if (zi || ((zmask >> q) & 1u)) g[q] = (i == q) ? 1.0 : 0.0;
      if (zi) diag0 = 1.0;
    }
    int bad = 0;
#pragma unroll
    for (int k = 0; k < PANEL_W; ++k) {
      double v = g[k];
#pragma unroll
      for (int q = 0; q < k; ++q) v -= g[q] * __shfl_sync(0xffffffffu, g[q], k);
      const double d = __shfl_sync(0xffffffffu, v, k), d0 = __shfl_sync(0xffffffffu, diag0, k);
      if (k < w && (!(d > CHOLQR_PIVOT_TOL * d0) || !(d0 > 0.0))) bad = 1;
      const double lkk = (d > 0.0) ? sqrt(d) : 1.0;
      g[k] = (i == k) ? lkk : v / lkk;
    }
    if (i < PANEL_W) {
#pragma unroll
      for (int q = 0; q < PANEL_W; ++q) G[i * PANEL_W + q] = g[q];   // L in the lower triangle (q <= i)
    }
    __syncwarp();
    if (!bad && i < w) {
      // column j = i of M = L^-1 (forward substitution); Ri[k][j] = R^-1[k][j] = M[j][k]
      const int j = i;
      double m[PANEL_W];
#pragma unroll
      for (int r = 0; r < PANEL_W; ++r) {
        double v = (r == j) ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < r; ++q) v -= G[r * PANEL_W + q] * ((q >= j) ? m[q] : 0.0);
        m[r] = (r >= j && r < w) ? v / G[r * PANEL_W + r] : 0.0;
      }
#pragma unroll
      for (int r = 0; r < PANEL_W; ++r) Ri[j * PANEL_W + r] = m[r];   // M[r][j] -> Ri[j][r]
    }
    if (i == 0) {
      *flag = bad;
      flag[1] = (int)zmask;
      if (!bad && zmask && jb0.nzero) *jb0.nzero += __popc(zmask);
    }
  }
#else
  // Cholesky G = L L^T (R = L^T) and Ri = R^-1 (simulator: one thread)
  PAR_FOR(one, 1) {
    int bad = 0;
    unsigned zm = 0u;
    for (int k = 0; k < w; ++k)
      if (!(G[k * PANEL_W + k] > rel_tol2 * jb0.norm0[k])) zm |= (1u << k);
    for (int k = 0; k < w; ++k)
      if ((zm >> k) & 1u)
        for (int q = 0; q < w; ++q) {
          G[k * PANEL_W + q] = (q == k) ? 1.0 : 0.0;
          G[q * PANEL_W + k] = (q == k) ? 1.0 : 0.0;
        }
    flag[1] = (int)zm;
    for (int k = 0; k < w && !bad; ++k) {
      double d = G[k * PANEL_W + k];
      const double d0 = d;
      for (int q = 0; q < k; ++q) d -= G[k * PANEL_W + q] * G[k * PANEL_W + q];   // L stored in the lower triangle
      if (!(d > CHOLQR_PIVOT_TOL * d0) || !(d0 > 0.0)) { bad = 1; break; }
      const double lkk = sqrt(d);
      G[k * PANEL_W + k] = lkk;
      for (int i = k + 1; i < w; ++i) {
        double v = G[i * PANEL_W + k];
        for (int q = 0; q < k; ++q) v -= G[i * PANEL_W + q] * G[k * PANEL_W + q];
        G[i * PANEL_W + k] = v / lkk;
      }
    }
    if (!bad) {
      // Ri = R^-1 with R[k][j] = L[j][k] (k <= j): back substitution column by column
      for (int j = 0; j < w; ++j) {
        for (int k = j; k >= 0; --k) {
          double v = (k == j) ? 1.0 : 0.0;
          for (int q = k + 1; q <= j; ++q) v -= G[q * PANEL_W + k] * Ri[q * PANEL_W + j];   // R[k][q] = L[q][k]
          Ri[k * PANEL_W + j] = v / G[k * PANEL_W + k];
        }
      }
    }
    *flag = bad;
    if (std::getenv("TMF_DEBUG_PANEL")) {
      double rmin = 1e300;
      for (int k = 0; k < w; ++k) rmin = std::min(rmin, G[k * PANEL_W + k]);
      std::fprintf(stderr, "[panel] rows %d c0 %d w %d bad %d zmask %x min L_kk %.3e\n", rows, c0, w, bad, zm, rmin);
    }
    if (!bad && zm && jb0.nzero) *jb0.nzero += __builtin_popcount(zm);
  }
#endif
  CTA_SYNC();
  if (*flag) {
    PanelJob jb = jb0;
    jb.use_smem = 0;
    CTA_SYNC();
    panel_mgs2_body(jb, rel_tol2, scratch);
    return;
  }
  // Q = P Ri, one row per thread; q_j only needs p_k with k <= j, so the row is transformed in place from
  // the last column down
  PAR_FOR(r, rows) {
    double p[PANEL_W];
    const unsigned zm = (unsigned)flag[1];
#pragma unroll
    for (int k = 0; k < PANEL_W; ++k) p[k] = (k < w && !((zm >> k) & 1u)) ? P[(int64_t)k * ld + r] : 0.0;
#pragma unroll
    for (int j = PANEL_W - 1; j >= 0; --j) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k <= j; ++k) v += p[k] * Ri[k * PANEL_W + j];
      p[j] = v;
    }
#pragma unroll
    for (int j = 0; j < PANEL_W; ++j)
      if (j < w) P[(int64_t)j * ld + r] = p[j];
  }
}
inline size_t panel_cholqr_smem_bytes(int c0_max = 0) {
  return sizeof(double) * (size_t)(2 * PANEL_W * PANEL_W + 2 + PANEL_W * 33 + PANEL_W + 40 + 8 + (size_t)c0_max * PANEL_W);
}
inline size_t panel_smem_bytes(int rows, int ncols, bool use_smem) {
  return sizeof(double) * (size_t)(PANEL_W * 33 + PANEL_W + 40 + (use_smem ? (size_t)rows * ncols : 0));
}

// ---------------------------------------------------------------------------------------------
// SVD of the small triangular factor: selects the entangled directions.
//   in : Rw (rr x rr, ld = rr)   out: Jsel (rr x rr): right singular vectors with
//        s^2 > thr first (k0 of them), remaining columns zero;  k0 -> *k0_out; status flag.
// ---------------------------------------------------------------------------------------------
struct SvdSelJob {
  const double *Rw;
  double *Jsel;
  double *Jwork;      // rr x rr global scratch for the rotations when they do not fit in smem
  int *k0_out;
  int *info;          // info[2] |= 1 when the sketch did not reach the noise floor
  const int *nzero;   // number of numerically dependent sketch columns (rank exhausted if > 0)
  int rr, complete;   // complete: the sketch spans the whole range of B (rr == min(n, m))
};
TMF_GLOBAL svd_select_kernel(const SvdSelJob *jobs, double thr, double floor2) {
  const SvdSelJob jb = jobs[BLOCK_ID];
  const int n = jb.rr;
  if (n <= 0) {
    PAR_FOR(one, 1) *jb.k0_out = 0;
    return;
  }
  DYN_SMEM(double, sm);
  double *G = sm, *s2 = G + n * n, *rot = s2 + n;
  double *part = rot + ((n + 1) & ~1);
  int *iw = reinterpret_cast<int *>(part + (((n + 1) & ~1) / 2) * 33 * 3 + 2);
  int *flag = iw, *sel = iw + 2, *rank = sel + n, *cnt = rank + n;
  // One-sided Jacobi on the *transposed* triangular factor (Drmac-Veselic): G = Rw^T, G J = U S,
  // so the normalised columns of the converged G are the right singular vectors of Rw, i.e. the
  // left singular vectors of W = Q^T B that we need; J itself is never formed.
  PAR_FOR(idx, n * n) {
    int c = idx / n, r = idx - c * n;
    G[idx] = jb.Rw[(size_t)r * n + c];
  }
  CTA_SYNC();
  const int sweeps = jacobi_onesided(G, n, nullptr, n, n, rot, part, flag);
  PAR_FOR(one, 1) jb.info[3] = sweeps;   // diagnostics: Jacobi sweeps of the sketch SVD
  PAR_FOR(c, n) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += G[c * n + r] * G[c * n + r];
    s2[c] = s;
    sel[c] = (s > thr) ? 1 : 0;
  }
  CTA_SYNC();
  rank_desc(s2, sel, n, rank, cnt);
  PAR_FOR(idx, n * n) jb.Jsel[idx] = 0.0;
  CTA_SYNC();
  PAR_FOR(idx, n * n) {
    int c = idx / n, r = idx - c * n;
    if (rank[c] >= 0) jb.Jsel[rank[c] * n + r] = G[idx] / sqrt(s2[c]);
  }
  PAR_FOR(one, 1) {
    *jb.k0_out = *cnt;
    // adequacy of the sketch: either the range of B was exhausted (dependent columns were
    // dropped) or the smallest captured singular value is far below the entanglement threshold.
    double smin = s2[0], smax = s2[0];
    for (int c = 1; c < n; ++c) { smin = fmin(smin, s2[c]); smax = fmax(smax, s2[c]); }
    bool ok = jb.complete || (jb.nzero && *jb.nzero > 0) || smin <= floor2 * smax || smax == 0.0;
    if (!ok) jb.info[2] |= 1;
  }
}
inline size_t svdsel_smem_bytes(int n) {
  int np = (n + 1) & ~1;
  size_t nj = 0;   // the SVD use needs no rotation matrix
  return sizeof(double) * ((size_t)n * n + nj + n + np + (size_t)(np / 2) * 33 * 3 + 2) +
         sizeof(int) * ((size_t)2 * n + 12);
}

// ---------------------------------------------------------------------------------------------
// Rayleigh-Ritz on the selected subspace: eigen-decomposition of TE (leading k0 x k0 block).
//   out: Zsel (rr x rr): eigenvectors of the entangled eigenvalues (cutoff <= e < 1-cutoff),
//        ordered by decreasing LEFT eigenvalue, zero-padded;  e_left -> e_out[0..k);
//        info[0] = k;  eside_out[0..k) = eigenvalue on this side (used by the Cholesky step).
// ---------------------------------------------------------------------------------------------
struct RitzJob {
  const double *TE;   // rr x rr, ld = rr
  double *Zsel;       // rr x rr
  double *Jwork;      // rr x rr global scratch (see SvdSelJob)
  const int *k0;
  double *e_out;      // TMF_MAX_MODES
  double *eside_out;  // TMF_MAX_MODES
  int *info;
  int rr, side;
};
TMF_GLOBAL ritz_kernel(const RitzJob *jobs, double cutoff) {
  const RitzJob jb = jobs[BLOCK_ID];
  const int n = jb.rr;
  const int k0 = (n > 0) ? *jb.k0 : 0;
  if (n <= 0 || k0 <= 0) {
    PAR_FOR(one, 1) jb.info[0] = 0;
    if (n > 0) { PAR_FOR(idx, n * n) jb.Zsel[idx] = 0.0; }
    return;
  }
  DYN_SMEM(double, sm);
  double *G = sm, *ev = G + n * n, *key = ev + n, *rot = key + n;
  double *part = rot + ((n + 1) & ~1);
  int *iw = reinterpret_cast<int *>(part + (((n + 1) & ~1) / 2) * 33 * 3 + 2);
  int *flag = iw, *sel = iw + 2, *rank = sel + n, *cnt = rank + n;
  double *J = (n > JAC_SMEM_J_MAX) ? jb.Jwork : reinterpret_cast<double *>(cnt + 6);
  PAR_FOR(idx, k0 * k0) {
    int c = idx / k0, r = idx - c * k0;
    // symmetrise the Rayleigh quotient matrix
    G[c * n + r] = 0.5 * (jb.TE[c * n + r] + jb.TE[r * n + c]);
    J[c * n + r] = (c == r) ? 1.0 : 0.0;
  }
  CTA_SYNC();
  jacobi_onesided(G, n, J, n, k0, rot, part, flag);
  // eigenvalues as Rayleigh quotients of the *original* matrix: the rotated columns of G carry the
  // accumulated rounding of all sweeps (~ sweeps * k * eps), the quotient only k * eps.
  PAR_FOR(idx, k0 * k0) {
    int c = idx / k0, r = idx - c * k0;
    double s = 0.0;
    for (int q = 0; q < k0; ++q) s += 0.5 * (jb.TE[q * n + r] + jb.TE[r * n + q]) * J[c * n + q];
    G[c * n + r] = s;
  }
  CTA_SYNC();
  PAR_FOR(c, k0) {
    double s = 0.0;
    for (int r = 0; r < k0; ++r) s += J[c * n + r] * G[c * n + r];
    double e = s;
    ev[c] = e;
    sel[c] = (e >= cutoff && e < 1.0 - cutoff) ? 1 : 0;
    key[c] = (jb.side == TMF_SIDE_L) ? e : 1.0 - e;  // left eigenvalue
  }
  CTA_SYNC();
  rank_desc(key, sel, k0, rank, cnt);
  PAR_FOR(idx, n * n) jb.Zsel[idx] = 0.0;
  CTA_SYNC();
  PAR_FOR(idx, k0 * k0) {
    int c = idx / k0, r = idx - c * k0;
    if (rank[c] >= 0) jb.Zsel[rank[c] * n + r] = J[c * n + r];
  }
  PAR_FOR(c, k0) {
    if (rank[c] >= 0 && rank[c] < TMF_MAX_MODES) {
      jb.e_out[rank[c]] = key[c];
      jb.eside_out[rank[c]] = ev[c];
    }
  }
  PAR_FOR(one, 1) {
    jb.info[0] = *cnt;
    if (*cnt > TMF_MAX_MODES) jb.info[2] |= 2;
  }
}

// ---------------------------------------------------------------------------------------------
// pivoted Cholesky of the filled-space projector P1 = A - U diag(w) U^T  (never formed).
// P1 is (numerically) an orthogonal projector, so its Cholesky factor has orthonormal columns:
// F (n x f) is an orthonormal basis of the eigenvalue-1 space of A.  Diagonal pivoting keeps
// every pivot >= rank/n.  One CTA per job; F and U stream from L2.
// ---------------------------------------------------------------------------------------------
struct CholJob {
  const double *A;     // n x n block of C, lda
  double *V;           // output matrix: columns [0,k) = U (input), [k, k+f) = F (output); ld = n
  const double *w;     // eigenvalues of U's columns on this side (k of them)
  int *info;           // reads info[0] = k, writes info[1] = f
  int n, lda, max_cols, pad_;
};
constexpr int PIVCHOL_B = 8;         // pivots per pass over the factor
constexpr int PIVCHOL_THREADS = 1024;

// Blocked left-looking variant.  The unblocked algorithm re-reads the whole factor [U | F] (n x K, in
// L2) for every pivot column -- sum_f n K 8 B ~ 34 GB at L = 1024, which made the kernel L2-bandwidth
// bound.  Here PIVCHOL_B candidate pivots (the largest residual diagonals) are advanced together: one
// pass over the factor produces all their columns (8 FMAs per L2 load), then the B x B coupling between
// them is resolved in shared memory.  A candidate whose residual diagonal collapses once its
// predecessors in the block are accounted for (it was nearly dependent on them) is skipped; its
// diagonal entry is exact, so the next selection simply does not pick it again.
TMF_GLOBAL pivchol_kernel(const CholJob *jobs, double tol) {
  const CholJob jb = jobs[BLOCK_ID];
  const int n = jb.n;
  if (n <= 0) return;
  const int k = jb.info[0];
  constexpr int B = PIVCHOL_B;
  DYN_SMEM(double, sm);
  double *d = sm;                        // n      residual diagonal (-1: already a pivot)
  double *dd = d + n;                    // n      scratch copy for the candidate selection
  double *cand = dd + n;                 // B * n  candidate columns, cand[b * n + i]
  double *lpb = cand + (size_t)B * n;    // B * n  rows p_b of the factor (scaled by w for U), lpb[m * B + b]
  double *psum = lpb + (size_t)B * n;    // B * PIVCHOL_THREADS partial sums
  double *red = psum + (size_t)B * PIVCHOL_THREADS;   // 72
  int *ired = reinterpret_cast<int *>(red + 72);      // 48: [0,32) lane results, [32, 32 + B) pivots, 44 count
  int *pidx = ired + 32;
  const double *U = jb.V;
  double *F = jb.V + (int64_t)k * n;
  PAR_FOR(i, n) {
    double s = jb.A[(int64_t)i * jb.lda + i];
    for (int m = 0; m < k; ++m) s -= jb.w[m] * U[(int64_t)m * n + i] * U[(int64_t)m * n + i];
    d[i] = s;
  }
  CTA_SYNC();
  int f = 0;
  const int fmax = jb.max_cols - k;
  bool done = false;
  while (f < fmax && !done) {
    // ---- candidate selection: the nb largest residual diagonals above tol ---------------------
    const int want = (fmax - f < B) ? (fmax - f) : B;
    PAR_FOR(i, n) dd[i] = d[i];
    CTA_SYNC();
    int nb = 0;
#if !defined(TMF_HOSTSIM)
    {
      // CTA-wide argmax per candidate with shuffles: thread-strided scan, warp reduction, one barrier,
      // then every warp reduces the 32 warp results redundantly (no second barrier, no single-thread
      // section).  The two-barrier / one-thread version below spent half of the kernel's time here.
      const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
      for (int b = 0; b < want; ++b) {
        double best = -1.0;
        int bi = 0x7fffffff;
        for (int i = tid; i < n; i += blockDim.x) {
          const double v = dd[i];
          if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        double *wr = psum + (b & 1) * 32;                          // psum is free during the selection
        int *wi = reinterpret_cast<int *>(psum + 64) + (b & 1) * 32;
        if (lane == 0) { wr[wid] = best; wi[wid] = bi; }
        __syncthreads();
        double fb = (lane < nw) ? wr[lane] : -1.0;
        int fi = (lane < nw) ? wi[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, fb, o);
          const int oi = __shfl_xor_sync(0xffffffffu, fi, o);
          if (ob > fb || (ob == fb && oi < fi)) { fb = ob; fi = oi; }
        }
        if (tid == 0) { red[40] = fb; red[48 + b] = fb; pidx[b] = fi; }
        if (!(fb > tol)) break;
        if (fi % (int)blockDim.x == tid) dd[fi] = -1.0;              // the owner re-reads it next round
        ++nb;
      }
      __syncthreads();
    }
#else
    for (int b = 0; b < want; ++b) {
      PAR_FOR(lane, 32) {
        double best = -1.0;
        int bi = 0;
        for (int i = lane; i < n; i += 32)
          if (dd[i] > best) { best = dd[i]; bi = i; }
        red[lane] = best;
        ired[lane] = bi;
      }
      CTA_SYNC();
      PAR_FOR(one, 1) {
        double best = red[0];
        int bi = ired[0];
        for (int l = 1; l < 32; ++l)
          if (red[l] > best || (red[l] == best && ired[l] < bi)) { best = red[l]; bi = ired[l]; }
        red[40] = best;
        red[48 + b] = best;      // residual diagonal at selection time
        pidx[b] = bi;
        if (best > tol) dd[bi] = -1.0;
      }
      CTA_SYNC();
      if (!(red[40] > tol)) break;
      ++nb;
    }
#endif
    if (nb == 0) break;
    const int K = k + f;
    // ---- rows p_b of the current factor --------------------------------------------------------
    PAR_FOR(item, K * B) {
      const int m = item / B, b = item - m * B;
      double v = 0.0;
      if (b < nb) {
        const int p = pidx[b];
        v = (m < k) ? jb.w[m] * U[(int64_t)m * n + p] : F[(int64_t)(m - k) * n + p];
      }
      lpb[item] = v;
    }
    CTA_SYNC();
    // ---- candidate columns: A[:, p_b] - [U | F] lp_b, one pass over the factor for all b -------
    {
      int parts = NTHREADS / ((n + 31) & ~31);
      parts = parts < 1 ? 1 : (parts > PIVCHOL_MAX_PARTS ? PIVCHOL_MAX_PARTS : parts);
      const int chunk = (K + parts - 1) / parts;
      PAR_FOR(item, n * parts) {
        const int part = item / n, i = item - part * n;
        const int m0 = part * chunk, m1 = (m0 + chunk < K) ? m0 + chunk : K;
        double acc[B];
#pragma unroll
        for (int b = 0; b < B; ++b) acc[b] = (part == 0 && b < nb) ? jb.A[(int64_t)pidx[b] * jb.lda + i] : 0.0;
        for (int m = m0; m < m1; ++m) {
          const double v = jb.V[(int64_t)m * n + i];
          const double *l = lpb + (size_t)m * B;
#pragma unroll
          for (int b = 0; b < B; ++b) acc[b] -= v * l[b];
        }
#pragma unroll
        for (int b = 0; b < B; ++b) psum[(size_t)item * B + b] = acc[b];
      }
      CTA_SYNC();
      PAR_FOR(item, n * nb) {
        const int b = item / n, i = item - b * n;
        double c = 0.0;
        for (int q = 0; q < parts; ++q) c += psum[(size_t)(q * n + i) * B + b];
        cand[(size_t)b * n + i] = c;
      }
      CTA_SYNC();
    }
    // ---- resolve the coupling inside the block ---------------------------------------------------
    for (int b = 0; b < nb && f < fmax; ++b) {
      const int p = pidx[b];
      const double piv = cand[(size_t)b * n + p];
      // a pivot that collapsed below a quarter of its value at selection time was nearly dependent on its
      // predecessors in this block: leave it to a later selection (d[p] already reflects the collapse)
      if (!(piv > tol) || piv < 0.25 * red[48 + b]) continue;
      const double inv = 1.0 / sqrt(piv);
      double *col = cand + (size_t)b * n;
      PAR_FOR(i, n) {
        const double l = col[i] * inv;
        col[i] = l;
        F[(int64_t)f * n + i] = l;
        d[i] = (i == p) ? -1.0 : d[i] - l * l;
      }
      CTA_SYNC();
      const int rest = nb - b - 1;
      if (rest > 0) {
        PAR_FOR(item, n * rest) {
          const int bb = b + 1 + item / n, i = item % n;
          cand[(size_t)bb * n + i] -= col[i] * col[pidx[bb]];
        }
        CTA_SYNC();
      }
      ++f;
    }
  }
  PAR_FOR(one, 1) jb.info[1] = f;
}
inline size_t pivchol_smem_bytes(int n) {
  return sizeof(double) * ((size_t)(2 + 2 * PIVCHOL_B) * n + (size_t)PIVCHOL_B * PIVCHOL_THREADS + 72) +
         sizeof(int) * 64;
}

// ---------------------------------------------------------------------------------------------
// direct path for small blocks (n <= SMALL_N): full Jacobi eigen-decomposition of A in shared
// memory, then split into entangled / filled exactly like slater.py:350-370.
// ---------------------------------------------------------------------------------------------
struct SmallJob {
  const double *A;
  double *V;       // n x n output, ld = n: [entangled (k) | filled (f)]
  double *e_out;   // TMF_MAX_MODES
  int *info;       // k, f, status, n
  int n, lda, side, pad_;
};
TMF_GLOBAL small_modes_kernel(const SmallJob *jobs, double cutoff) {
  const SmallJob jb = jobs[BLOCK_ID];
  const int n = jb.n;
  if (n <= 0) {
    PAR_FOR(one, 1) { jb.info[0] = 0; jb.info[1] = 0; }
    return;
  }
  DYN_SMEM(double, sm);
  double *G = sm, *J = G + n * n, *ev = J + n * n, *key = ev + n, *rot = key + n;
  double *part = rot + ((n + 1) & ~1);
  int *iw = reinterpret_cast<int *>(part + (((n + 1) & ~1) / 2) * 33 * 3 + 2);
  int *flag = iw, *sel = iw + 2, *rank = sel + n, *cnt = rank + n, *fil = cnt + 2, *frank = fil + n;
  PAR_FOR(idx, n * n) {
    int c = idx / n, r = idx - c * n;
    G[idx] = 0.5 * (jb.A[(int64_t)c * jb.lda + r] + jb.A[(int64_t)r * jb.lda + c]);
    J[idx] = (c == r) ? 1.0 : 0.0;
  }
  CTA_SYNC();
  jacobi_onesided(G, n, J, n, n, rot, part, flag);
  // eigenvalues as Rayleigh quotients of the original block (see ritz_kernel)
  PAR_FOR(idx, n * n) {
    int c = idx / n, r = idx - c * n;
    double s = 0.0;
    for (int q = 0; q < n; ++q)
      s += 0.5 * (jb.A[(int64_t)q * jb.lda + r] + jb.A[(int64_t)r * jb.lda + q]) * J[c * n + q];
    G[idx] = s;
  }
  CTA_SYNC();
  PAR_FOR(c, n) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += J[c * n + r] * G[c * n + r];
    double e = s;
    ev[c] = e;
    sel[c] = (e >= cutoff && e < 1.0 - cutoff) ? 1 : 0;
    fil[c] = (e >= 1.0 - cutoff) ? 1 : 0;
    key[c] = (jb.side == TMF_SIDE_L) ? e : 1.0 - e;
  }
  CTA_SYNC();
  rank_desc(key, sel, n, rank, cnt);
  const int k = *cnt;
  CTA_SYNC();
  rank_desc(ev, fil, n, frank, cnt);
  const int f = *cnt;
  PAR_FOR(idx, n * n) {
    int c = idx / n, r = idx - c * n;
    int dst = rank[c] >= 0 ? rank[c] : (frank[c] >= 0 ? k + frank[c] : -1);
    if (dst >= 0) jb.V[(int64_t)dst * n + r] = J[idx];
  }
  PAR_FOR(c, n)
    if (rank[c] >= 0 && rank[c] < TMF_MAX_MODES) jb.e_out[rank[c]] = key[c];
  PAR_FOR(one, 1) {
    jb.info[0] = k;
    jb.info[1] = f;
    if (k > TMF_MAX_MODES) jb.info[2] |= 2;
  }
}
// ---------------------------------------------------------------------------------------------
// Edge vector of the filled space (nested-projector site stage, see siteprep.cu).
// The filled space F of the block A gains at most one dimension when the block grows by one site, and
// the new direction is the projection of that site's unit vector: g = P_F e_edge / |P_F e_edge| with
// P_F = A - U diag(lam) U^T (never formed).  Written to column k of V (right after the k entangled
// modes), orthogonalised against them.  Also derives the number of filled orbitals from the trace:
// f = round(tr A - sum lam)  (the modes that are not selected contribute < cutoff each).
// ---------------------------------------------------------------------------------------------
struct EdgeJob {
  const double *A;       // n x n diagonal block of C (symmetric), lda
  double *V;             // column-major, ld = n: columns [0,k) entangled modes (in), column k = g (out)
  const double *e_left;  // left eigenvalues of the k modes
  int *info;             // reads info[0] = k, writes info[1] = f
  double *edge_out;      // [0] = |P_F e_edge|^2, [1] = tr A - sum lam - f
  int n, lda, side, emb;   // emb = 1: re/im-interleaved embedding of a complex matrix (the edge site has two rows)
};
TMF_GLOBAL edge_vector_kernel(const EdgeJob *jobs) {
  const EdgeJob jb = jobs[BLOCK_ID];
  const int n = jb.n;
  if (n <= 0) {
    PAR_FOR(one, 1) { jb.info[1] = 0; jb.edge_out[0] = 0.0; jb.edge_out[1] = 0.0; }
    return;
  }
  int k = jb.info[0];
  if (k > TMF_MAX_MODES) k = TMF_MAX_MODES;   // flagged by the Ritz kernel (status 2)
  const int edge = (jb.side == TMF_SIDE_R) ? 0 : n - (jb.emb ? 2 : 1);   // (embedded: the row of the real part)
  DYN_SMEM(double, sm);
  double *lam = sm, *wx = lam + TMF_MAX_MODES, *coef = wx + TMF_MAX_MODES;
  double *part = coef + TMF_MAX_MODES;          // (TMF_MAX_MODES + 2) * 33
  double *g = jb.V + (int64_t)k * n;
  PAR_FOR(j, k) {
    const double l = (jb.side == TMF_SIDE_L) ? jb.e_left[j] : 1.0 - jb.e_left[j];
    lam[j] = l;
    wx[j] = l * jb.V[(int64_t)j * n + edge];
  }
  PAR_FOR(lane, 32) {
    double t = 0.0;
    for (int r = lane; r < n; r += 32) t += jb.A[(int64_t)r * jb.lda + r];
    part[TMF_MAX_MODES * 33 + lane] = t;
  }
  CTA_SYNC();
  PAR_FOR(r, n) {
    double v = jb.A[(int64_t)edge * jb.lda + r];
    for (int j = 0; j < k; ++j) v -= jb.V[(int64_t)j * n + r] * wx[j];
    g[r] = v;
  }
  CTA_SYNC();
  PAR_FOR(item, k * 32) {
    const int j = item >> 5, lane = item & 31;
    const double *u = jb.V + (int64_t)j * n;
    double t = 0.0;
    for (int r = lane; r < n; r += 32) t += u[r] * g[r];
    part[j * 33 + lane] = t;
  }
  CTA_SYNC();
  PAR_FOR(j, k) {
    double t = 0.0;
    for (int l = 0; l < 32; ++l) t += part[j * 33 + l];
    coef[j] = t;
  }
  CTA_SYNC();
  PAR_FOR(r, n) {
    double v = g[r];
    for (int j = 0; j < k; ++j) v -= jb.V[(int64_t)j * n + r] * coef[j];
    g[r] = v;
  }
  CTA_SYNC();
  PAR_FOR(lane, 32) {
    double t = 0.0;
    for (int r = lane; r < n; r += 32) t += g[r] * g[r];
    part[(TMF_MAX_MODES + 1) * 33 + lane] = t;
  }
  CTA_SYNC();
  double gn2 = 0.0, tr = 0.0;
  for (int l = 0; l < 32; ++l) { gn2 += part[(TMF_MAX_MODES + 1) * 33 + l]; tr += part[TMF_MAX_MODES * 33 + l]; }
  const double inv = (gn2 > 0.0) ? 1.0 / sqrt(gn2) : 0.0;
  CTA_SYNC();
  PAR_FOR(r, n) g[r] *= inv;
  PAR_FOR(one, 1) {
    double s = tr;
    for (int j = 0; j < k; ++j) s -= lam[j];
    const double f = floor(s + 0.5);
    jb.info[1] = (int)f;
    jb.edge_out[0] = gn2;
    jb.edge_out[1] = s - f;
  }
}
inline size_t edge_smem_bytes() { return sizeof(double) * (3 * TMF_MAX_MODES + (TMF_MAX_MODES + 2) * 33); }

// ---------------------------------------------------------------------------------------------
// Complex modes out of the real eigenvectors of an embedded Hermitian matrix (complex Slater path).
// The re/im-interleaved embedding of a complex vector w is the real vector emb(w); every complex eigenvector
// of the Hermitian block appears in the embedded real block as the plane span{emb(w), J emb(w)} (J = times i)
// with a doubly degenerate eigenvalue, and the real solver returns an arbitrary orthonormal basis of it (or of
// the union of several planes when eigenvalues coincide numerically).  Any unit vector of the plane is emb of
// a phase times w -- a valid eigenvector -- so the k complex modes are picked by a complex Gram-Schmidt sweep
// over the 2k real columns in eigenvalue order: a column whose remainder after removing the *complex* span of
// the modes chosen so far keeps more than a quarter of its norm opens a new mode.  In place: chosen mode j
// (as interleaved complex = the same memory as a real column) goes to column j, its eigenvalue to e[j], the
// edge vector moves from column 2k to column k.
// ---------------------------------------------------------------------------------------------
struct PairCJob {
  double *V;        // n_emb x (k_emb + 1), ld = n_emb
  double *e_left;   // in: k_emb eigenvalues (pairs), out: k
  int *info;        // in: [0] = k_emb, [1] = f_emb; out: k, f; [2] |= 4 on failure
  int n_emb, side;
};
TMF_GLOBAL pair_complex_kernel(const PairCJob *jobs) {
  const PairCJob jb = jobs[BLOCK_ID];
  const int n = jb.n_emb, ke = jb.info[0] > TMF_MAX_MODES ? TMF_MAX_MODES : jb.info[0];
  DYN_SMEM(double, sm);
  double *part = sm;                     // 3 * 33
  double *eo = part + 3 * 33 + 1;        // TMF_MAX_MODES: eigenvalues of the chosen modes
  double *absw = eo + TMF_MAX_MODES;     // weight of the skipped columns inside every chosen mode's plane
  double *tmpw = absw + TMF_MAX_MODES;
  int *misc = reinterpret_cast<int *>(tmpw + TMF_MAX_MODES);
  if (n <= 0) return;
  PAR_FOR(j, TMF_MAX_MODES) absw[j] = 0.0;
  PAR_FOR(one, 1) misc[0] = 0;
  CTA_SYNC();
  for (int c = 0; c < ke; ++c) {
    double *a = jb.V + (int64_t)c * n;
    const int m = misc[0];
    for (int j = 0; j < m; ++j) {
      const double *w = jb.V + (int64_t)j * n;
      // coef = w^dagger a  (complex inner product of the interleaved vectors)
      PAR_FOR(lane, 32) {
        double re = 0.0, im = 0.0;
        for (int r = 2 * lane; r < n; r += 64) {
          re += w[r] * a[r] + w[r + 1] * a[r + 1];
          im += w[r] * a[r + 1] - w[r + 1] * a[r];
        }
        part[lane] = re; part[33 + lane] = im;
      }
      CTA_SYNC();
      double re = 0.0, im = 0.0;
      for (int l = 0; l < 32; ++l) { re += part[l]; im += part[33 + l]; }
      CTA_SYNC();
      PAR_FOR(one, 1) tmpw[j] = re * re + im * im;
      PAR_FOR(h, n / 2) {
        const double wr = w[2 * h], wi = w[2 * h + 1];
        a[2 * h] -= wr * re - wi * im;
        a[2 * h + 1] -= wr * im + wi * re;
      }
      CTA_SYNC();
    }
    PAR_FOR(lane, 32) {
      double t = 0.0;
      for (int r = lane; r < n; r += 32) t += a[r] * a[r];
      part[66 + lane] = t;
    }
    CTA_SYNC();
    double nrm2 = 0.0;
    for (int l = 0; l < 32; ++l) nrm2 += part[66 + l];
    CTA_SYNC();
    if (nrm2 > 0.25 && m < TMF_MAX_MODES / 2) {
      const double inv = 1.0 / sqrt(nrm2);
      double *dst = jb.V + (int64_t)m * n;
      PAR_FOR(r, n) dst[r] = a[r] * inv;
      PAR_FOR(one, 1) { eo[m] = jb.e_left[c]; misc[0] = m + 1; }
    } else {
      PAR_FOR(j, m) absw[j] += tmpw[j];
    }
    CTA_SYNC();
  }
  const int k = misc[0];
  // edge vector: column k_emb -> column k (k <= k_emb; equal only for k_emb <= 1, when nothing has to move),
  // orthogonalised against the *complex* span of the chosen modes (a no-op unless a pair was split by the
  // entanglement cutoff: then its lone member stands for the complex mode and the partner J v leaves the filled space)
  double *g = jb.V + (int64_t)k * n;
  if (k != ke) {
    const double *src = jb.V + (int64_t)ke * n;
    PAR_FOR(r, n) g[r] = src[r];
    CTA_SYNC();
  }
  for (int j = 0; j < k; ++j) {
    const double *w = jb.V + (int64_t)j * n;
    PAR_FOR(lane, 32) {
      double re = 0.0, im = 0.0;
      for (int r = 2 * lane; r < n; r += 64) {
        re += w[r] * g[r] + w[r + 1] * g[r + 1];
        im += w[r] * g[r + 1] - w[r + 1] * g[r];
      }
      part[lane] = re; part[33 + lane] = im;
    }
    CTA_SYNC();
    double re = 0.0, im = 0.0;
    for (int l = 0; l < 32; ++l) { re += part[l]; im += part[33 + l]; }
    CTA_SYNC();
    PAR_FOR(h, n / 2) {
      const double wr = w[2 * h], wi = w[2 * h + 1];
      g[2 * h] -= wr * re - wi * im;
      g[2 * h + 1] -= wr * im + wi * re;
    }
    CTA_SYNC();
  }
  PAR_FOR(lane, 32) {
    double t = 0.0;
    for (int r = lane; r < n; r += 32) t += g[r] * g[r];
    part[66 + lane] = t;
  }
  CTA_SYNC();
  double gn2 = 0.0;
  for (int l = 0; l < 32; ++l) gn2 += part[66 + l];
  CTA_SYNC();
  if (gn2 > 0.0) {
    const double inv = 1.0 / sqrt(gn2);
    PAR_FOR(r, n) g[r] *= inv;
  }
  PAR_FOR(one, 1) {
    // lone members (no skipped partner inside their plane): the partner was classified filled (side eigenvalue
    // ~ 1) or empty (~ 0) by the cutoff; the complex mode counts as entangled, the filled count loses the partner
    int f_emb = jb.info[1], lone = 0;
    for (int j = 0; j < k; ++j)
      if (absw[j] < 0.5) {
        ++lone;
        const double lam = (jb.side == TMF_SIDE_L) ? eo[j] : 1.0 - eo[j];
        if (lam > 0.5) --f_emb;
      }
#if defined(TMF_HOSTSIM)
    if ((2 * k - lone != ke || (f_emb & 1) || f_emb < 0) && std::getenv("TMF_DEBUG_PAIR"))
      fprintf(stderr, "pair_complex: n_emb %d k_emb %d chosen %d lone %d f_emb %d -> %d\n", n, ke, k, lone, jb.info[1], f_emb);
#endif
    if (2 * k - lone != ke || (f_emb & 1) || f_emb < 0) jb.info[2] |= 4;
    jb.info[0] = k;
    jb.info[1] = f_emb / 2;
  }
  CTA_SYNC();
  PAR_FOR(j, TMF_MAX_MODES) jb.e_left[j] = (j < k) ? eo[j] : 0.0;
}
inline size_t pairc_smem_bytes() { return sizeof(double) * (3 * 33 + 1 + 3 * TMF_MAX_MODES) + 64; }

inline size_t ritz_smem_bytes(int n) {
  int np = (n + 1) & ~1;
  size_t nj = (n > JAC_SMEM_J_MAX) ? 0 : (size_t)n * n;
  return sizeof(double) * ((size_t)n * n + nj + 2 * n + np + (size_t)(np / 2) * 33 * 3 + 2) +
         sizeof(int) * ((size_t)2 * n + 12);
}
inline size_t small_smem_bytes(int n) {
  int np = (n + 1) & ~1;
  return sizeof(double) * ((size_t)2 * n * n + 2 * n + np + (size_t)(np / 2) * 33 * 3 + 2) +
         sizeof(int) * ((size_t)4 * n + 12);
}

}  // namespace tmf
