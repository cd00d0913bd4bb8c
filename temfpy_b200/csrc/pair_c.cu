// K4 for complex Slater determinants: pairing of the left and right entangled modes of the central bond
// (utils.block_svd, utils.py:19-96, as called from slater.py:407, and the sign flips of :410) with complex mode
// matrices.  C_LR V_R runs on the real DMMA GEMM through the embedding (emb(A) emb(v) = emb(A v)); the k x k
// matrix M = V_L^H C_LR V_R and the rotations of the k columns are small complex kernels; the SVDs of the
// degenerate k x k groups are done on the host (one-sided complex Jacobi).
#include <cmath>
#include "cplx.hpp"

namespace tmf {
int gemm_grouped(const tmf_gemm_job *jobs, int njobs, void *desc_dev, void *stream);
int64_t gemm_desc_bytes(int njobs);

// out (ka x kb complex, row-major) = A^H B for interleaved complex column-major A (n x ka), B (n x kb)
TMF_GLOBAL cdot_block_kernel(const double *A, int lda, int ka, const double *B, int ldb, int kb, int n, double *out) {
  const cplx *a = reinterpret_cast<const cplx *>(A), *b = reinterpret_cast<const cplx *>(B);
  cplx *o = reinterpret_cast<cplx *>(out);
  DYN_SMEM(double, part);   // 2 * 33 per output handled in a pass
  const int i = BLOCK_ID / kb, j = BLOCK_ID % kb;
  if (i >= ka) return;
  PAR_FOR(lane, 32) {
    cplx s = cmake(0.0);
    for (int r = lane; r < n; r += 32) s = cadd(s, cmulc(a[(int64_t)i * lda + r], b[(int64_t)j * ldb + r]));
    part[lane] = s.x; part[33 + lane] = s.y;
  }
  CTA_SYNC();
  PAR_FOR(one, 1) {
    cplx s = cmake(0.0);
    for (int l = 0; l < 32; ++l) s = cadd(s, cmake(part[l], part[33 + l]));
    o[i * kb + j] = s;
  }
}

// tmp (n x k) = V[:, :k] R  (R: k x k complex, column-major: R[j + i * k] = entry (j, i)); then V[:, :k] <- tmp
TMF_GLOBAL crotate_kernel(double *V, int ld, int n, int k, const double *R, double *tmp) {
  const cplx *v = reinterpret_cast<const cplx *>(V), *rot = reinterpret_cast<const cplx *>(R);
  cplx *t = reinterpret_cast<cplx *>(tmp);
  const int i = BLOCK_ID;       // output column
  PAR_FOR(r, n) {
    cplx s = cmake(0.0);
    for (int j = 0; j < k; ++j) s = cadd(s, cmul(v[(int64_t)j * ld + r], rot[(int64_t)i * k + j]));
    t[(int64_t)i * n + r] = s;
  }
}
TMF_GLOBAL ccopy_cols_kernel(double *V, int ld, int n, const double *tmp) {
  cplx *v = reinterpret_cast<cplx *>(V);
  const cplx *t = reinterpret_cast<const cplx *>(tmp);
  const int i = BLOCK_ID;
  PAR_FOR(r, n) v[(int64_t)i * ld + r] = t[(int64_t)i * n + r];
}

namespace {
// one-sided complex Jacobi SVD of an m x m matrix G (column-major): G = U diag(s) V^H
void small_svd_c(std::vector<cplx> G, int m, std::vector<cplx> &U, std::vector<cplx> &V) {
  V.assign((size_t)m * m, cmake(0.0));
  for (int i = 0; i < m; ++i) V[(size_t)i * m + i] = cmake(1.0);
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool any = false;
    for (int p = 0; p < m; ++p)
      for (int q = p + 1; q < m; ++q) {
        double a = 0, b = 0;
        cplx c = cmake(0.0);
        for (int r = 0; r < m; ++r) {
          a += cabs2(G[p * m + r]);
          b += cabs2(G[q * m + r]);
          c = cadd(c, cmulc(G[p * m + r], G[q * m + r]));    // g_p^H g_q
        }
        const double ac = std::sqrt(cabs2(c));
        if (ac == 0.0 || ac <= 1e-15 * std::sqrt(a) * std::sqrt(b)) continue;
        any = true;
        const cplx ph = cscale(c, 1.0 / ac);                 // phase of the off-diagonal Gram entry
        const double zeta = (b - a) / (2 * ac);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
        const double cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
        // columns: p' = cs p - sn conj(ph) q,  q' = sn ph p + cs q   (unitary, annihilates g_p'^H g_q')
        for (int r = 0; r < m; ++r) {
          const cplx x = G[p * m + r], y = G[q * m + r];
          G[p * m + r] = csub(cscale(x, cs), cscale(cmul(cconj(ph), y), sn));
          G[q * m + r] = cadd(cscale(cmul(ph, x), sn), cscale(y, cs));
          const cplx vx = V[p * m + r], vy = V[q * m + r];
          V[p * m + r] = csub(cscale(vx, cs), cscale(cmul(cconj(ph), vy), sn));
          V[q * m + r] = cadd(cscale(cmul(ph, vx), sn), cscale(vy, cs));
        }
      }
    if (!any) break;
  }
  U.assign((size_t)m * m, cmake(0.0));
  for (int c = 0; c < m; ++c) {
    double s = 0;
    for (int r = 0; r < m; ++r) s += cabs2(G[c * m + r]);
    s = std::sqrt(s);
    for (int r = 0; r < m; ++r) U[c * m + r] = (s > 0) ? cscale(G[c * m + r], 1.0 / s) : cmake(r == c ? 1.0 : 0.0);
  }
}
}  // namespace
}  // namespace tmf

extern "C" int64_t tmf_slater_pair_bond_c_workspace(int L, int k) {
  return tmf::align256(16 * (int64_t)L * (k + 1)) * 3 + tmf::align256(16 * (int64_t)k * k) * 3 +
         tmf::gemm_desc_bytes(4) + 4096;
}

// Complex form of tmf_slater_pair_bond.  Cemb_dev: the 2L x 2L real embedding (pitch ldc doubles); VL (x rows) /
// VR (L - x rows): interleaved complex mode matrices (ld = rows, complex elements); their first k columns are
// rotated in place.  Synchronises once (k x k matrix to the host).
extern "C" int tmf_slater_pair_bond_c(const double *Cemb_dev, int ldc, int L, int x, int k, const double *e_host,
                                      double degeneracy_tol, double *VL, double *VR, void *work_dev,
                                      int64_t work_bytes, void *stream) {
  using namespace tmf;
  if (k <= 0) return TMF_OK;
  if (x <= 0 || x >= L) { set_error("tmf_slater_pair_bond_c: bond has an empty side"); return TMF_ERR_VALUE; }
  Arena ar(work_dev, work_bytes);
  const int nL = x, nR = L - x;
  double *T1 = ar.take<double>(2 * (int64_t)nL * k), *tmp = ar.take<double>(2 * (int64_t)std::max(nL, nR) * k);
  double *M = ar.take<double>(2 * (int64_t)k * k), *RotL = ar.take<double>(2 * (int64_t)k * k);
  double *RotR = ar.take<double>(2 * (int64_t)k * k);
  void *desc = ar.take<unsigned char>(gemm_desc_bytes(4));
  if (!ar.ok()) { set_error("workspace too small (complex pairing)"); return TMF_ERR_VALUE; }
  // T1 = emb(C_LR) VR: rows 0..2x of columns 2x.. of the embedded matrix (symmetric: read as k-contiguous rows)
  tmf_gemm_job j1;
  std::memset(&j1, 0, sizeof(j1));
  j1.A = Cemb_dev + 2 * (int64_t)x; j1.lda = ldc; j1.transA = 1;       // op(A)[m, kk] = Cemb[m, 2x + kk]
  j1.B = VR; j1.ldb = 2 * nR; j1.transB = 0;
  j1.C = T1; j1.ldc = 2 * nL;
  j1.M = 2 * nL; j1.N = k; j1.K = 2 * nR; j1.alpha = 1.0; j1.beta = 0.0;
  int rc = gemm_grouped(&j1, 1, desc, stream);
  if (rc) return rc;
  rc = launch_t("pair_c", cdot_block_kernel, k * k, 32, sizeof(double) * 70, stream, (const double *)VL, nL, k,
                (const double *)T1, nL, k, nL, M);
  if (rc) return rc;
  std::vector<cplx> Mh((size_t)k * k);
  rc = copy_d2h_sync(Mh.data(), M, sizeof(cplx) * Mh.size(), stream);     // Mh[i * k + j] = (VL^H C_LR VR)[i][j]
  if (rc) return rc;
  std::vector<cplx> RL((size_t)k * k, cmake(0.0)), RR((size_t)k * k, cmake(0.0));
  int a = 0;
  while (a < k) {  // groups of (nearly) degenerate eigenvalues (utils.py:71-78)
    int b = a + 1;
    while (b < k && !(std::fabs(e_host[b] - e_host[b - 1]) > degeneracy_tol)) ++b;
    const int m = b - a;
    std::vector<cplx> G((size_t)m * m), U, V;
    // column-major m x m block: G[cc * m + r] = M[a + r][a + cc]
    for (int cc = 0; cc < m; ++cc)
      for (int r = 0; r < m; ++r) G[(size_t)cc * m + r] = Mh[(size_t)(a + r) * k + a + cc];
    small_svd_c(G, m, U, V);       // block = U S V^H: vL <- vL U, vR <- vR V   (utils.py:90-94)
    for (int cc = 0; cc < m; ++cc)
      for (int r = 0; r < m; ++r) {
        RL[(size_t)(a + cc) * k + a + r] = U[(size_t)cc * m + r];
        RR[(size_t)(a + cc) * k + a + r] = V[(size_t)cc * m + r];
      }
    a = b;
  }
  for (int i = 0; i < k; ++i)     // slater.py:410: reference column j of vRE is mode k-1-j; odd j flips
    if ((k - 1 - i) & 1)
      for (int r = 0; r < k; ++r) RR[(size_t)i * k + r] = cneg(RR[(size_t)i * k + r]);
  rc = copy_h2d(RotL, RL.data(), sizeof(cplx) * RL.size(), stream);
  if (rc) return rc;
  rc = copy_h2d(RotR, RR.data(), sizeof(cplx) * RR.size(), stream);
  if (rc) return rc;
  rc = launch_t("pair_c", crotate_kernel, k, 256, 0, stream, VL, nL, nL, k, (const double *)RotL, tmp);
  if (rc) return rc;
  rc = launch_t("pair_c", ccopy_cols_kernel, k, 256, 0, stream, VL, nL, nL, (const double *)tmp);
  if (rc) return rc;
  rc = launch_t("pair_c", crotate_kernel, k, 256, 0, stream, VR, nR, nR, k, (const double *)RotR, tmp);
  if (rc) return rc;
  return launch_t("pair_c", ccopy_cols_kernel, k, 256, 0, stream, VR, nR, nR, (const double *)tmp);
}
