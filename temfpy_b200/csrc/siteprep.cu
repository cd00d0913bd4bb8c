// K8 + K9: overlap of the mode bases of two neighbouring bonds and Schur complement.
//
// reference: slater.py:1071  O = HT(v_bra) @ v_ket   (after _select_orbitals reordered / signed
// the columns, :1066-1067) and :1073-1090 (det of the always-always block, sometimes matrix
// D - C A^-1 B).  The explicit inverse of the reference is replaced by a blocked LU of the
// k x k always block whose elimination is carried through the remaining rows and columns: after
// k steps the trailing (rows-k) x (cols-k) block *is* the Schur complement and the product of the
// pivots is det_always.  Pivoting is restricted to the rows of the always block.
//
// O is produced by the grouped DMMA GEMM (column gathers + sign scalings in the epilogue); the LU
// runs one CTA per site with 16-wide panels staged in shared memory.
#include <cstdlib>
#include "cta.hpp"

namespace tmf {

int gemm_launch_uploaded(const tmf_gemm_job *jobs_dev, const int *prefix_dev, int njobs, int ntiles,
                         void *stream, const char *tag);

constexpr int NB = 16;
constexpr double SCHUR_MIN_PIVOT = 1e-9;
static_assert(sizeof(tmf_site_job) == 128, "site descriptor must be 128 bytes");

// Frame of the elimination: the side with more always-occupied orbitals provides the rows, so that
// the k = min(ka_bra, ka_ket) pivots can be chosen among all of its always orbitals.
struct Frame {
  int tr;            // 1: rows = ket orbitals, cols = bra orbitals (O is stored transposed)
  int R, Cc;         // frame rows / cols
  int ncand, k;      // candidate pivot rows, elimination steps
  int s_r, s_c;      // sometimes orbitals of the row / column side
};
TMF_HD Frame make_frame(const tmf_site_job &jb) {
  Frame f;
  f.tr = jb.ka_ket > jb.ka_bra;
  const int ka_r = f.tr ? jb.ka_ket : jb.ka_bra, ka_c = f.tr ? jb.ka_bra : jb.ka_ket;
  f.s_r = f.tr ? jb.sk : jb.sb;
  f.s_c = f.tr ? jb.sb : jb.sk;
  f.R = ka_r + f.s_r;
  f.Cc = ka_c + f.s_c;
  f.ncand = ka_r;
  f.k = ka_c;
  return f;
}

TMF_GLOBAL schur_kernel(const tmf_site_job *jobs) {
  const tmf_site_job jb = jobs[BLOCK_ID];
  const Frame fr = make_frame(jb);
  const int rows = fr.R, cols = fr.Cc, k = fr.k, ld = fr.R;
  double *O = jb.O;
  DYN_SMEM(double, sm);
  double *Lp = sm;                      // rows x NB  (column-major, ld = nr of the current panel)
  double *Up = Lp + (size_t)rows * NB;  // NB x cols  (row-major, ld = nc)
  double *red = Up + (size_t)cols * NB; // 40
  int *ired = reinterpret_cast<int *>(red + 40);  // 40 + NB
  int *piv = ired + 40;
  if (rows <= 0 || cols <= 0) {
    if (jb.pad_[0] == 0) { PAR_FOR(one, 1) *jb.det = 1.0; }
    return;
  }
  // physical orbital: <n_i| overlaps = one row of the ket mode matrix (slater.py:1030-1051); in the
  // embedded Pfaffian frame (emb) the two physical modes (c^+ / c rows of pfaffian.py:1667-1688) are
  // fixed combinations of the site's four Majorana components
  const bool nested = jb.pad_[0] != 0;   // O and the running determinant were prepared by nested_site_kernel
  if (jb.physical && !nested) {
    const int unit = jb.emb ? 4 : 1;
    const int src = ((jb.mode == 1) ? 0 : jb.n_bra) * unit;
    const int nb = jb.ka_bra + jb.sb, nk = jb.ka_ket + jb.sk;
    const double h = 0.70710678118654752440;
    for (int pb = 0; pb < nb; ++pb) {
      const int code = jb.bra_cols[pb];
      if (code >= 0) continue;
      double cf[4] = {1.0, 0.0, 0.0, 0.0};
      if (jb.emb) {
        const int t = -1 - code;   // 0: emb(lo) 1: J emb(lo) 2: emb(up) 3: J emb(up)
        cf[0] = (t == 0 || t == 2) ? h : 0.0;
        cf[1] = (t == 1 || t == 3) ? h : 0.0;
        cf[2] = (t == 1) ? -h : (t == 3 ? h : 0.0);
        cf[3] = (t == 0) ? h : (t == 2 ? -h : 0.0);
      }
      PAR_FOR(n, nk) {
        int c = jb.ket_cols[n];
        double v = 0.0;
        if (c >= 0) {
          const double *col = jb.Vk + (int64_t)c * jb.ldk + src;
          v = cf[0] * col[0];
          if (jb.emb) v += cf[1] * col[1] + cf[2] * col[2] + cf[3] * col[3];
        }
        v *= jb.bra_sign[pb] * jb.ket_sign[n];
        if (fr.tr) O[(int64_t)pb * ld + n] = v;
        else O[(int64_t)n * ld + pb] = v;
      }
    }
  }
  PAR_FOR(one, 1) {
    red[36] = nested ? *jb.det : 1.0;  // running determinant
    red[38] = 1.0;                     // smallest |pivot|
  }
  CTA_SYNC();

  for (int t0 = 0; t0 < k; t0 += NB) {
    const int w = (k - t0 < NB) ? (k - t0) : NB;
    const int nr = rows - t0;         // panel rows t0 .. rows-1
    const int lim = fr.ncand - t0;    // rows eligible as pivots (always orbitals of the row side)
    PAR_FOR(idx, nr * w) {
      int c = idx / nr, r = idx - c * nr;
      Lp[c * nr + r] = O[(int64_t)(t0 + c) * ld + t0 + r];
    }
    CTA_SYNC();
    for (int j = 0; j < w; ++j) {
      PAR_FOR(lane, 32) {
        double best = -1.0;
        int bi = j;
        for (int r = j + lane; r < lim; r += 32) {
          double a = fabs(Lp[j * nr + r]);
          if (a > best) { best = a; bi = r; }
        }
        red[lane] = best;
        ired[lane] = bi;
      }
      CTA_SYNC();
      PAR_FOR(one, 1) {
        double best = red[0];
        int bi = ired[0];
        for (int l = 1; l < 32; ++l)
          if (red[l] > best || (red[l] == best && ired[l] < bi)) { best = red[l]; bi = ired[l]; }
        piv[j] = bi;
        double pv = Lp[j * nr + bi];
        red[36] *= (bi != j) ? -pv : pv;
        red[37] = (pv != 0.0) ? 1.0 / pv : 0.0;
        red[38] = fmin(red[38], fabs(pv));
      }
      CTA_SYNC();
      const int p = piv[j];
      if (p != j) {
        PAR_FOR(c, w) {
          double a = Lp[c * nr + j];
          Lp[c * nr + j] = Lp[c * nr + p];
          Lp[c * nr + p] = a;
        }
        CTA_SYNC();
      }
      const double inv = red[37];
      PAR_FOR(r, nr - j - 1) Lp[j * nr + j + 1 + r] *= inv;
      CTA_SYNC();
      const int hr = nr - j - 1;
      PAR_FOR(idx, (w - j - 1) * hr) {
        int c = j + 1 + idx / hr, r = j + 1 + idx % hr;
        Lp[c * nr + r] -= Lp[j * nr + r] * Lp[c * nr + j];
      }
      CTA_SYNC();
    }
    PAR_FOR(idx, nr * w) {
      int c = idx / nr, r = idx - c * nr;
      O[(int64_t)(t0 + c) * ld + t0 + r] = Lp[c * nr + r];
    }
    const int nc = cols - t0 - w;
    // row interchanges + forward substitution on the block row, one thread per column
    PAR_FOR(jc, nc) {
      double *col = O + (int64_t)(t0 + w + jc) * ld + t0;
      for (int j = 0; j < w; ++j) {
        int p = piv[j];
        if (p != j) { double a = col[j]; col[j] = col[p]; col[p] = a; }
      }
      double u[NB];
      for (int a = 0; a < w; ++a) {
        double v = col[a];
        for (int b = 0; b < a; ++b) v -= Lp[b * nr + a] * u[b];
        u[a] = v;
      }
      for (int a = 0; a < w; ++a) {
        col[a] = u[a];
        Up[a * nc + jc] = u[a];
      }
    }
    CTA_SYNC();
    // trailing update with 4 x 4 register tiles
    const int mr = nr - w;
    const int ti = (mr + 3) / 4, tj = (nc + 3) / 4;
    PAR_FOR(tile, ti * tj) {
      int i0 = (tile % ti) * 4, j0 = (tile / ti) * 4;
      double acc[4][4];
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
      for (int kk = 0; kk < w; ++kk) {
        double l[4], u[4];
        for (int a = 0; a < 4; ++a) l[a] = (i0 + a < mr) ? Lp[kk * nr + w + i0 + a] : 0.0;
        for (int b = 0; b < 4; ++b) u[b] = (j0 + b < nc) ? Up[kk * nc + j0 + b] : 0.0;
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) acc[a][b] += l[a] * u[b];
      }
      for (int b = 0; b < 4; ++b)
        for (int a = 0; a < 4; ++a)
          if (i0 + a < mr && j0 + b < nc)
            O[(int64_t)(t0 + w + j0 + b) * ld + t0 + w + i0 + a] -= acc[a][b];
    }
    CTA_SYNC();
  }
  // Schur complement -> S in the reference's row / column order (slater.py:1081-1090): the surplus
  // always orbitals of the row side come first (left tensors) or last (right tensors).
  const int nlo = fr.ncand - k, nfr = nlo + fr.s_r, nfc = fr.s_c;
  const int s_bra = fr.tr ? nfc : nfr;
  PAR_FOR(idx, nfr * nfc) {
    int c = idx / nfr, r = idx - c * nfr;
    int ref_r = (jb.mode == 0) ? r : (r < nlo ? fr.s_r + r : r - nlo);
    double v = O[(int64_t)(k + c) * ld + k + r];
    if (fr.tr) jb.S[(int64_t)ref_r * s_bra + c] = v;   // row side = ket -> column of S
    else jb.S[(int64_t)c * s_bra + ref_r] = v;
  }
  // A vanishing pivot = an always-occupied orbital of one bond that is orthogonal to every always-occupied
  // orbital of the other: the kept Schmidt vectors of the two bonds are incompatible (a numerically degenerate
  // multiplet of spin-pure modes was cut differently on the two bonds).  Reported as NaN; the driver redoes the
  // conversion with symmetrised weights (TMF_OPT_SNAP).  Only for jobs that ask for it (pad_[1], chain driver).
  PAR_FOR(one, 1) *jb.det = (jb.pad_[1] != 0 && red[38] < SCHUR_MIN_PIVOT) ? nan("") : red[36];
}


// ---------------------------------------------------------------------------------------------
// Nested-projector site stage (chain driver): the sometimes matrix without any basis of the filled
// spaces.
//
// The reference takes O = V_bra^T V_ket over filled + entangled orbitals and eliminates the always
// occupied block (slater.py:1071-1090).  The filled spaces of two neighbouring blocks are nested up to
// the truncation threshold: F_ket = F_bra (+) g (the edge vector of edge_vector_kernel) when the filled
// count grows, F_ket = F_bra otherwise -- exactly for eigenvalue-1 orbitals, and for the thresholded
// spaces (1 - e < cutoff) up to "leaks" of F_bra into the weakly entangled ket modes.  With
//   Y_b = [e_site | X_bra]   (physical orbital + entangled bra modes),   Y_k = [X_ket | g],
//   P_b = A_b - X_b E_b X_b^T,  P_k' = A_k - Y_k Lam' Y_k^T   (filled-space projectors, never formed),
//   Z_k = P_b Y_k  (leak of the ket orbitals into F_bra),  Z_b = P_k' Y_b,
// the Schur complement over the filled spaces is, exactly up to O(cutoff),
//   S = Y_b^T Y_k - (Z_b^T Z_k) (1 - Z_k^T Z_k)^-1,      |det_always| = sqrt(det(1 - Z_k^T Z_k)),
// and by the eigen-relations A_b X_b = X_b E_b, A_k Y_k = Y_k Lam' every matrix above is a small
// polynomial in  T = X_b^T Y_k',  w = X_b^T a,  u = Y_k'^T a,  a^T a,  y0 = Y_k[edge, :]  (a = the column
// of C that couples the new site to the bra block, ' = rows of the bra block).  One CTA per site:
// one pass over the two n x k mode matrices (instead of the n x n filled bases, a 277^2 GEMM and a
// k = 255 LU per site), then (k+1)^3 work in shared memory.  The result is written in the frame
// schur_kernel expects; that kernel then eliminates the always-occupied *entangled* orbitals.
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(tmf_nested_job) == 64, "nested descriptor must be 64 bytes");
constexpr int NS_RC = 32;   // rows per staged chunk

TMF_GLOBAL nested_site_kernel(const tmf_site_job *jobs, const tmf_nested_job *njobs) {
  const tmf_site_job jb = jobs[BLOCK_ID];
  const tmf_nested_job nj = njobs[BLOCK_ID];
  const Frame fr = make_frame(jb);
  const int kb = nj.k_bra, kk = nj.k_ket, ck = kk + nj.df, rb = kb + 1;
  const int n_b = jb.n_bra, n_k = n_b + 1;
  const int off = (jb.mode == 1) ? 1 : 0, edge = (jb.mode == 1) ? 0 : n_k - 1;
  const int ldq = ck + 1;
  DYN_SMEM(double, sm);
  double *Q = sm;                              // (kb + 1) x (ck + 1): rows X_b | a, cols Y_k' | a
  double *y0 = Q + (size_t)(kb + 1) * ldq;     // ck
  double *lamk = y0 + (ck + 1);                // ck
  double *eb = lamk + (ck + 1);                // kb
  double *aZ = eb + (kb + 1);                  // ck
  double *red = aZ + (ck + 1);                 // 4
  double *reg = red + 4;                       // staging / phase-2 matrices
  double *XZ = reg;                            // kb x ck
  double *YZ = XZ + (size_t)kb * ck;           // ck x ck
  double *H = YZ + (size_t)ck * ck;            // ck x ck
  double *K = H + (size_t)ck * ck;             // rb x ck
  const int nbc = kb + 1, ncol = nbc + ck, lds = NS_RC + 1;
  double *st = reg;                            // ncol x (NS_RC + 1)

  PAR_FOR(idx, (kb + 1) * ldq) Q[idx] = 0.0;
  CTA_SYNC();
  const int ntr = (kb + 2) / 2, ntc = (ck + 2) / 2;   // 2 x 2 output tiles over (kb + 1) x (ck + 1)
  for (int r0 = 0; r0 < n_b; r0 += NS_RC) {
    PAR_FOR(idx, ncol * NS_RC) {
      const int col = idx / NS_RC, r = idx - col * NS_RC;
      const int gr = r0 + r;
      double v = 0.0;
      if (gr < n_b) {
        if (col < kb) v = jb.Vb[(int64_t)col * jb.ldb + gr];
        else if (col == kb) v = nj.a_col[gr];
        else v = jb.Vk[(int64_t)(col - nbc) * jb.ldk + gr + off];
      }
      st[col * lds + r] = v;
    }
    CTA_SYNC();
    PAR_FOR(tile, ntr * ntc) {
      const int i0 = (tile % ntr) * 2, j0 = (tile / ntr) * 2;
      const int i1 = (i0 + 1 < kb + 1) ? i0 + 1 : i0, j1 = (j0 + 1 < ck + 1) ? j0 + 1 : j0;
      // ket-side column ck is the coupling column a, staged once as bra column kb
      const double *b0 = st + (size_t)i0 * lds, *b1 = st + (size_t)i1 * lds;
      const double *k0 = st + (size_t)(j0 < ck ? nbc + j0 : kb) * lds, *k1 = st + (size_t)(j1 < ck ? nbc + j1 : kb) * lds;
      double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
      for (int r = 0; r < NS_RC; ++r) {
        const double x0 = b0[r], x1 = b1[r], z0 = k0[r], z1 = k1[r];
        a00 += x0 * z0; a01 += x0 * z1; a10 += x1 * z0; a11 += x1 * z1;
      }
      Q[i0 * ldq + j0] += a00;
      if (j1 != j0) Q[i0 * ldq + j1] += a01;
      if (i1 != i0) Q[i1 * ldq + j0] += a10;
      if (i1 != i0 && j1 != j0) Q[i1 * ldq + j1] += a11;
    }
    CTA_SYNC();
  }
  // small vectors
  const int side = (jb.mode == 1) ? TMF_SIDE_R : TMF_SIDE_L;
  PAR_FOR(j, ck) {
    y0[j] = jb.Vk[(int64_t)j * jb.ldk + edge];
    lamk[j] = (j < kk) ? ((side == TMF_SIDE_L) ? nj.e_ket[j] : 1.0 - nj.e_ket[j]) : 1.0;
  }
  PAR_FOR(i, kb) eb[i] = (side == TMF_SIDE_L) ? nj.e_bra[i] : 1.0 - nj.e_bra[i];
  CTA_SYNC();
  const double aa = Q[kb * ldq + ck];
  // X_b^T Z_k, a^T Z_k, Y'^T Z_k
  PAR_FOR(idx, kb * ck) {
    const int i = idx / ck, j = idx - i * ck;
    XZ[idx] = Q[i * ldq + j] * (lamk[j] - eb[i]) - Q[i * ldq + ck] * y0[j];
  }
  PAR_FOR(j, ck) {
    double v = Q[kb * ldq + j] * lamk[j] - aa * y0[j];
    for (int i = 0; i < kb; ++i) v -= Q[i * ldq + ck] * eb[i] * Q[i * ldq + j];
    aZ[j] = v;
  }
  PAR_FOR(idx, ck * ck) {
    const int j = idx / ck, l = idx - j * ck;
    double v = ((j == l ? 1.0 : 0.0) - y0[j] * y0[l]) * lamk[l] - Q[kb * ldq + j] * y0[l];
    for (int i = 0; i < kb; ++i) v -= Q[i * ldq + j] * eb[i] * Q[i * ldq + l];
    YZ[idx] = v;
  }
  CTA_SYNC();
  // G = Z_k^T Z_k (symmetrised into H = 1 - G),  K = Z_b^T Z_k
  PAR_FOR(idx, ck * ck) {
    const int j = idx / ck, l = idx - j * ck;
    double v = lamk[j] * YZ[idx] - y0[j] * aZ[l];
    for (int i = 0; i < kb; ++i) v -= Q[i * ldq + j] * eb[i] * XZ[i * ck + l];
    H[idx] = v;
  }
  PAR_FOR(idx, rb * ck) {
    const int r = idx / ck, l = idx - r * ck;
    double v;
    if (r == 0) {
      v = aZ[l];
      for (int j = 0; j < ck; ++j) v -= y0[j] * lamk[j] * YZ[j * ck + l];
    } else {
      const int i = r - 1;
      v = eb[i] * XZ[i * ck + l];
      for (int j = 0; j < ck; ++j) v -= Q[i * ldq + j] * lamk[j] * YZ[j * ck + l];
    }
    K[idx] = v;
  }
  CTA_SYNC();
  // YZ <- H = 1 - (G + G^T) / 2   (lower triangle used)
  PAR_FOR(idx, ck * ck) {
    const int j = idx / ck, l = idx - j * ck;
    YZ[idx] = (j == l ? 1.0 : 0.0) - 0.5 * (H[j * ck + l] + H[l * ck + j]);
  }
  PAR_FOR(one, 1) { red[0] = 1.0; red[1] = 0.0; }
  CTA_SYNC();
  double *Lc = YZ;   // Cholesky factor, lower, row-major
  for (int j = 0; j < ck; ++j) {
    PAR_FOR(one, 1) {
      const double d = Lc[j * ck + j];
      // cos^2 of a principal angle between the two filled spaces: far from 1 only if they are not nested
      if (!(d > 1e-3)) red[1] = 1.0;
      const double sd = sqrt(d > 1e-300 ? d : 1e-300);
      Lc[j * ck + j] = sd;
      red[0] *= sd;
    }
    CTA_SYNC();
    const double inv = 1.0 / Lc[j * ck + j];
    PAR_FOR(i, ck - j - 1) Lc[(j + 1 + i) * ck + j] *= inv;
    CTA_SYNC();
    const int m = ck - j - 1;
    PAR_FOR(idx, m * m) {
      const int i = j + 1 + idx / m, l = j + 1 + idx % m;
      if (l <= i) Lc[i * ck + l] -= Lc[i * ck + j] * Lc[l * ck + j];
    }
    CTA_SYNC();
  }
  // rows of K <- S0 - K H^-1   (H = L L^T)
  PAR_FOR(r, rb) {
    double *x = K + (size_t)r * ck;
    for (int j = 0; j < ck; ++j) {
      double v = x[j];
      for (int l = 0; l < j; ++l) v -= Lc[j * ck + l] * x[l];
      x[j] = v / Lc[j * ck + j];
    }
    for (int j = ck - 1; j >= 0; --j) {
      double v = x[j];
      for (int l = j + 1; l < ck; ++l) v -= Lc[l * ck + j] * x[l];
      x[j] = v / Lc[j * ck + j];
    }
    for (int j = 0; j < ck; ++j) x[j] = ((r == 0) ? y0[j] : Q[(r - 1) * ldq + j]) - x[j];
  }
  CTA_SYNC();
  // gather into the elimination frame of schur_kernel (orbital order and signs of the site plan)
  const int nbo = jb.ka_bra + jb.sb, nko = jb.ka_ket + jb.sk, ld = fr.R;
  PAR_FOR(idx, nbo * nko) {
    const int pb = idx % nbo, n = idx / nbo;
    const int cb = jb.bra_cols[pb], cn = jb.ket_cols[n];
    const int rbn = (cb < 0) ? 0 : 1 + cb;
    const double v = K[(size_t)rbn * ck + cn] * jb.bra_sign[pb] * jb.ket_sign[n];
    if (fr.tr) jb.O[(int64_t)pb * ld + n] = v;
    else jb.O[(int64_t)n * ld + pb] = v;
  }
  PAR_FOR(one, 1) *jb.det = (red[1] != 0.0) ? nan("") : red[0];
}

static size_t nested_smem_bytes(int kb, int ck) {
  const size_t head = (size_t)(kb + 1) * (ck + 1) + 3 * (size_t)(ck + 1) + (kb + 1) + 4;
  const size_t phase2 = (size_t)kb * ck + 2 * (size_t)ck * ck + (size_t)(kb + 1) * ck;
  const size_t stage = (size_t)(kb + 1 + ck) * (NS_RC + 1);
  return sizeof(double) * (head + std::max(phase2, stage) + 8);
}

static size_t schur_smem_bytes(int rows, int cols) {
  return sizeof(double) * ((size_t)(rows + cols) * NB + 40) + sizeof(int) * (40 + NB + 8);
}

}  // namespace tmf

extern "C" int tmf_site_overlap_schur_batched(const tmf_site_job *jobs_host, int nsites,
                                              void *desc_dev, void *stream) {
  using namespace tmf;
  if (nsites <= 0) return TMF_OK;
  // descriptor layout in desc_dev: site jobs | gemm jobs | gemm prefix
  std::vector<tmf_gemm_job> g(nsites);
  std::vector<int> prefix(nsites + 1, 0);
  size_t smem = 0;
  for (int s = 0; s < nsites; ++s) {
    const tmf_site_job &sj = jobs_host[s];
    const Frame fr = make_frame(sj);
    tmf_gemm_job &j = g[s];
    std::memset(&j, 0, sizeof(j));
    const int unit = sj.emb ? 4 : 1;                         // rows per site (Pfaffian frame: 4)
    const int off = (sj.physical && sj.mode == 1) ? unit : 0;  // right mode: ket site 0 is the new site
    if (!fr.tr) {
      j.A = sj.Vb; j.lda = sj.ldb; j.a_idx = sj.bra_cols; j.row_scale = sj.bra_sign; j.a_row_off = 0;
      j.B = sj.Vk; j.ldb = sj.ldk; j.b_idx = sj.ket_cols; j.col_scale = sj.ket_sign; j.b_row_off = off;
    } else {
      j.A = sj.Vk; j.lda = sj.ldk; j.a_idx = sj.ket_cols; j.row_scale = sj.ket_sign; j.a_row_off = off;
      j.B = sj.Vb; j.ldb = sj.ldb; j.b_idx = sj.bra_cols; j.col_scale = sj.bra_sign; j.b_row_off = 0;
    }
    j.C = sj.O;
    j.M = fr.R; j.N = fr.Cc; j.K = sj.n_bra * unit;
    j.ldc = fr.R;
    j.transA = 1; j.transB = 0;
    j.alpha = 1.0; j.beta = 0.0;
    int tm = (j.M + 63) / 64, tn = (j.N + 63) / 64;
    if (j.M <= 0 || j.N <= 0) tm = tn = 0;
    prefix[s + 1] = prefix[s] + tm * tn;
    smem = std::max(smem, schur_smem_bytes(fr.R, fr.Cc));
  }
  if (smem > 220 * 1024) {
    set_error("site too large for the shared-memory LU panels (rows + cols > ~1700)");
    return TMF_ERR_VALUE;
  }
  unsigned char *d = static_cast<unsigned char *>(desc_dev);
  const size_t o_site = 0, o_gemm = align256(sizeof(tmf_site_job) * (size_t)nsites);
  const size_t o_pref = o_gemm + align256(sizeof(tmf_gemm_job) * (size_t)nsites);
  int rc = copy_h2d(d + o_site, jobs_host, sizeof(tmf_site_job) * (size_t)nsites, stream);
  if (rc) return rc;
  rc = copy_h2d(d + o_gemm, g.data(), sizeof(tmf_gemm_job) * (size_t)nsites, stream);
  if (rc) return rc;
  rc = copy_h2d(d + o_pref, prefix.data(), sizeof(int) * (size_t)(nsites + 1), stream);
  if (rc) return rc;
  rc = gemm_launch_uploaded(reinterpret_cast<const tmf_gemm_job *>(d + o_gemm),
                            reinterpret_cast<const int *>(d + o_pref), nsites, prefix[nsites], stream, "gemm_site");
  if (rc) return rc;
  static const int schur_threads = std::getenv("TMF_SCHUR_THREADS") ? std::atoi(std::getenv("TMF_SCHUR_THREADS")) : 256;
  return launch_t("schur", schur_kernel, nsites, schur_threads, smem, stream,
                reinterpret_cast<const tmf_site_job *>(d + o_site));
}

extern "C" int64_t tmf_site_desc_bytes(int nsites) {
  return tmf::align256(128 * (int64_t)nsites) * 2 + tmf::align256(4 * (int64_t)(nsites + 1)) + 256;
}

// Nested-projector form of the call above for the chain driver: `njobs_host[s]` carries what the
// fused kernel needs beyond the site job (eigenvalues, the coupling column of C, mode counts).
extern "C" int tmf_site_nested_batched(const tmf_site_job *jobs_host, const tmf_nested_job *njobs_host,
                                       int nsites, void *desc_dev, void *stream) {
  using namespace tmf;
  if (nsites <= 0) return TMF_OK;
  size_t smem_n = 0, smem_s = 0;
  std::vector<tmf_site_job> sj(jobs_host, jobs_host + nsites);
  for (int s = 0; s < nsites; ++s) {
    const tmf_nested_job &nj = njobs_host[s];
    if (nj.df < 0 || nj.df > 1 || nj.k_bra < 0 || nj.k_ket < 0 || nj.k_bra > TMF_MAX_MODES || nj.k_ket > TMF_MAX_MODES ||
        !sj[s].physical || sj[s].emb) {
      set_error("tmf_site_nested_batched: filled spaces of neighbouring bonds are not nested (df not in {0, 1})");
      return TMF_ERR_VALUE;
    }
    sj[s].pad_[0] = 1;
    const Frame fr = make_frame(sj[s]);
    smem_n = std::max(smem_n, nested_smem_bytes(nj.k_bra, nj.k_ket + nj.df));
    smem_s = std::max(smem_s, schur_smem_bytes(fr.R, fr.Cc));
  }
  unsigned char *d = static_cast<unsigned char *>(desc_dev);
  const size_t o_site = 0, o_nest = align256(sizeof(tmf_site_job) * (size_t)nsites);
  int rc = copy_h2d(d + o_site, sj.data(), sizeof(tmf_site_job) * (size_t)nsites, stream);
  if (rc) return rc;
  rc = copy_h2d(d + o_nest, njobs_host, sizeof(tmf_nested_job) * (size_t)nsites, stream);
  if (rc) return rc;
  rc = launch_t("nested_site", nested_site_kernel, nsites, 256, smem_n, stream,
                reinterpret_cast<const tmf_site_job *>(d + o_site), reinterpret_cast<const tmf_nested_job *>(d + o_nest));
  if (rc) return rc;
  return launch_t("schur", schur_kernel, nsites, 128, smem_s, stream, reinterpret_cast<const tmf_site_job *>(d + o_site));
}
