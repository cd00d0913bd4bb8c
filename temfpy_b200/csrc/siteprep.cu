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
    PAR_FOR(one, 1) *jb.det = 1.0;
    return;
  }
  // physical orbital: <n_i| overlaps = one row of the ket mode matrix (slater.py:1030-1051); in the
  // embedded Pfaffian frame (emb) the two physical modes (c^+ / c rows of pfaffian.py:1667-1688) are
  // fixed combinations of the site's four Majorana components
  if (jb.physical) {
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
  PAR_FOR(one, 1) red[36] = 1.0;  // running determinant
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
  PAR_FOR(one, 1) *jb.det = red[36];
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
