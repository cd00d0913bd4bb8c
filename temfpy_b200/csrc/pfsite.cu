// K9p: per-site finish of the Pfaffian site stage on the device (reference pfaffian.py:1339-1400).
//
// tmf_site_overlap_schur_batched has eliminated the non-entangled modes of both bonds; what is left per site is a
// real matrix R of the size of the entangled modes (rows: [surplus | upper bra pairs | lower bra pairs], columns:
// [surplus | upper ket pairs | lower ket pairs], every complex number a 2 x 2 real block).  The reference now takes
// the square "upper-upper" block X = U*, its singular values (norm of the overlap of the two vacua; a vanishing one
// means opposite vacuum parity, pfaffian.py:1352-1357), its inverse (:1384) and assembles the antisymmetric
// contraction matrix N = [[BB, BA], [-BA^T, AA]] of :1386-1400 on the modes that are occupied in any Schmidt vector.
// One CTA per site: the sign / swap fixes of :1665, :1708-1719, :915-916 and the centre-bond rotations of block_svd
// (:855) are applied while the matrix is loaded, X is diagonalised by the one-sided Jacobi routine (X J = U Sigma,
// hence sigma, and X^-1 = J Sigma^-2 (U Sigma)^T without a separate factorisation), and N is written where
// tmf_pfaffians_blocks reads it.  The driver launches it twice: first for the singular values only (the vacuum
// parities of the whole chain follow from which overlaps vanish), then with the fixes that depend on them.
#include "jacobi.cuh"

namespace tmf {

static_assert(sizeof(tmf_pf_site_job) == 128, "descriptor must be 128 bytes");
constexpr int PFS_MAX = 80;     // largest dimension of R (4 (k + 1) + surplus, k <= 16)
constexpr int PFS_NX = 44;      // largest dimension of X (2 (k + 1) + surplus)

TMF_GLOBAL pf_site_kernel(const tmf_pf_site_job *jobs) {
  const tmf_pf_site_job jb = jobs[BLOCK_ID];
  const int sb = jb.sb, sk = jb.sk, sur_b = jb.sur_b, sur_k = jb.sur_k, mode = jb.mode;
  const int a1 = jb.k1 + (jb.no_phys ? 0 : 1), a2 = jb.k2;   // bra modes incl. the physical one (absent: iMPS gauge overlap)
  const int nr = sb + sur_b, nc = sk + sur_k;
  const int nx = sur_b + 2 * a1;                       // X is nx x nx (== sur_k + 2 a2)
  DYN_SMEM(double, sm);
  double *R = sm;                                      // nr x nc, column-major
  double *G = R + PFS_MAX * PFS_MAX;                   // nx x nx: X, then X J
  double *J = G + PFS_NX * PFS_NX;                     // nx x nx
  double *Xi = J + PFS_NX * PFS_NX;                    // nx x nx: X^-1
  double *tmp = G;                                     // G | J | Xi as one scratch (nr x 2 a2) for the rotations
  double *sig = Xi + PFS_NX * PFS_NX;                  // PFS_MAX
  double *rot = sig + PFS_MAX;                         // jacobi scratch: PFS_MAX + 2
  double *part = rot + PFS_MAX + 2;                    // (PFS_MAX / 2) * 99 + 40
  int *flag = reinterpret_cast<int *>(part + (PFS_MAX / 2) * 99 + 40);
  int *lst = flag + 4;                                 // index lists R1 | U1 | U2 | C2, each <= 2 * 17
  if (nr > PFS_MAX || nc > PFS_MAX || nx > PFS_NX || nx != sur_k + 2 * a2 || nx > nr || nx > nc ||
      (int64_t)nr * 2 * a2 > 3 * PFS_NX * PFS_NX) {
    PAR_FOR(one, 1) { jb.out[0] = -1.0; jb.out[1] = -1.0; }     // inconsistent mode counts
    return;
  }
  // ---- load with the surplus rows / columns in front (right tensors store them last) -------------------------
  PAR_FOR(idx, nr * nc) {
    const int c = idx / nr, r = idx - c * nr;
    const int ro = (mode == 1) ? (r < sur_b ? sb + r : r - sur_b) : r;
    const int co = (mode == 1) ? (c < sur_k ? sk + c : c - sur_k) : c;
    R[c * nr + r] = jb.S[(int64_t)co * nr + ro];
  }
  CTA_SYNC();
  // ---- centre-bond rotations of the ket modes (upper and lower pairs) -----------------------------------------
  if (jb.rot_up != nullptr && a2 > 0) {
    const int w = 2 * a2;
    for (int h = 0; h < 2; ++h) {
      const double *Q = h ? jb.rot_lo : jb.rot_up;     // w x w, row-major: new[:, j] = sum_i old[:, i] Q[i, j]
      const int base = sur_k + h * w;
      PAR_FOR(idx, nr * w) {
        const int j = idx / nr, r = idx - j * nr;
        double s = 0.0;
        for (int i = 0; i < w; ++i) s += R[(base + i) * nr + r] * Q[i * w + j];
        tmp[j * nr + r] = s;
      }
      CTA_SYNC();
      PAR_FOR(idx, nr * w) {
        const int j = idx / nr, r = idx - j * nr;
        R[(base + j) * nr + r] = tmp[j * nr + r];
      }
      CTA_SYNC();
    }
  }
  // ---- signs and the physical-mode swap ---------------------------------------------------------------------------
  const int phys = (mode == 1) ? 0 : a1 - 1;     // the physical mode, or the most entangled bra mode (pfaffian.py:1709-1719)
  const int up0 = sur_b, lo0 = sur_b + 2 * a1;
  const int p0 = up0 + 2 * phys, p1 = lo0 + 2 * phys;   // first rows of the upper / lower pair of the physical mode
  PAR_FOR(idx, nr * nc) {
    const int c = idx / nr, r = idx - c * nr;
    const bool is_phys = (r == p0 || r == p0 + 1 || r == p1 || r == p1 + 1);
    double f = 1.0;
    if (is_phys) f *= jb.u_p;
    if (c >= sur_k) f *= jb.ket_sign;
    if (jb.fix && mode == 1 && !is_phys && r >= up0 && r < lo0 + 2 * a1) f = -f;
    if (f != 1.0) R[idx] *= f;
  }
  CTA_SYNC();
  if (jb.fix) {
    PAR_FOR(idx, 2 * nc) {
      const int c = idx >> 1, t = idx & 1;
      const double a = R[c * nr + p0 + t], b = R[c * nr + p1 + t];
      R[c * nr + p0 + t] = b;
      R[c * nr + p1 + t] = a;
    }
    CTA_SYNC();
  }
  // ---- X = R[:nx, :nx]: singular values by one-sided Jacobi -------------------------------------------------------
  if (nx == 0) {
    PAR_FOR(one, 1) { jb.out[0] = 1.0; jb.out[1] = 1e300; }
  } else {
    PAR_FOR(idx, nx * nx) {
      const int c = idx / nx, r = idx - c * nx;
      G[idx] = R[c * nr + r];
      J[idx] = (r == c) ? 1.0 : 0.0;
    }
    CTA_SYNC();
    PAR_FOR(lane, 32) {
      double s = 0.0;
      for (int i = lane; i < nx * nx; i += 32) s += G[i] * G[i];
      part[lane] = s;
    }
    CTA_SYNC();
    double fro = 0.0;
    for (int l = 0; l < 32; ++l) fro += part[l];
    fro = sqrt(fro);
    CTA_SYNC();
    if (fro > 0.0) {
      const double rf = 1.0 / fro;
      PAR_FOR(idx, nx * nx) G[idx] *= rf;
      CTA_SYNC();
      jacobi_onesided(G, nx, J, nx, nx, rot, part, flag);
      PAR_FOR(idx, nx * nx) G[idx] *= fro;
      CTA_SYNC();
    }
    PAR_FOR(i, nx) {
      double s = 0.0;
      for (int r = 0; r < nx; ++r) s += G[i * nx + r] * G[i * nx + r];
      sig[i] = sqrt(s);
    }
    CTA_SYNC();
    PAR_FOR(one, 1) {
      double pr = 1.0, mn = 1e300;
      for (int i = 0; i < nx; ++i) { pr *= sig[i]; mn = fmin(mn, sig[i]); }
      jb.out[0] = pr;
      jb.out[1] = mn;
    }
  }
  if (!jb.want_n) return;
  // ---- X^-1 = J Sigma^-2 (X J)^T ----------------------------------------------------------------------------------
  PAR_FOR(idx, nx * nx) {
    const int j = idx / nx, i = idx - j * nx;          // Xi[i, j], column-major
    double s = 0.0;
    for (int t = 0; t < nx; ++t) {
      const double sg = sig[t];
      if (sg > 0.0) s += J[t * nx + i] * G[t * nx + j] / (sg * sg);
    }
    Xi[j * nx + i] = s;
  }
  // index lists: active bra modes ascending, active ket modes descending (pfaffian.py:1361-1374)
  PAR_FOR(one, 1) {
    int n1 = 0, n2 = 0;
    for (int t = 0; t < a1; ++t)
      if ((jb.idx1_mask >> t) & 1u) {
        lst[2 * n1] = lo0 + 2 * t; lst[2 * n1 + 1] = lo0 + 2 * t + 1;                  // R1
        lst[40 + 2 * n1] = up0 + 2 * t; lst[40 + 2 * n1 + 1] = up0 + 2 * t + 1;        // U1
        ++n1;
      }
    for (int t = a2 - 1; t >= 0; --t)
      if ((jb.idx2_mask >> t) & 1u) {
        lst[80 + 2 * n2] = sur_k + 2 * t; lst[80 + 2 * n2 + 1] = sur_k + 2 * t + 1;    // U2
        lst[120 + 2 * n2] = sur_k + 2 * a2 + 2 * t; lst[120 + 2 * n2 + 1] = sur_k + 2 * a2 + 2 * t + 1;   // C2
        ++n2;
      }
    flag[2] = n1;
    flag[3] = n2;
  }
  CTA_SYNC();
  const int n1 = flag[2], n2 = flag[3], m = n1 + n2;
  const int *R1 = lst, *U1 = lst + 40, *U2 = lst + 80, *C2 = lst + 120;
  // complex entry (a, b) of a J-structured real product: real part [2a, 2b], imaginary part [2a+1, 2b]
  // AA = R[R1, :nx] Xi[:, U1];  BA = Xi[U2, U1];  BB = Xi[U2, :] R[:nx, C2];  N = [[BB, BA], [-BA^T, AA]]
  // (AA and BB antisymmetrised), written complex, row-major, m x m
  double *AAr = G, *BBr = J;                           // (n1 x n1) and (n2 x n2) complex, interleaved
  PAR_FOR(idx, n1 * n1 * 2) {
    const int t = idx & 1, ab = idx >> 1, a = ab / n1, b = ab - a * n1;
    const int row = R1[2 * a + t], col = U1[2 * b];
    double s = 0.0;
    for (int u = 0; u < nx; ++u) s += R[u * nr + row] * Xi[col * nx + u];
    AAr[idx] = s;
  }
  PAR_FOR(idx, n2 * n2 * 2) {
    const int t = idx & 1, ab = idx >> 1, a = ab / n2, b = ab - a * n2;
    const int row = U2[2 * a + t], col = C2[2 * b];
    double s = 0.0;
    for (int u = 0; u < nx; ++u) s += Xi[u * nx + row] * R[col * nr + u];
    BBr[idx] = s;
  }
  CTA_SYNC();
  PAR_FOR(idx, m * m * 2) {
    const int t = idx & 1, ij = idx >> 1, i = ij / m, j = ij - i * m;
    double v;
    if (i < n2 && j < n2) {
      v = 0.5 * (BBr[2 * (i * n2 + j) + t] - BBr[2 * (j * n2 + i) + t]);
    } else if (i >= n2 && j >= n2) {
      const int a = i - n2, b = j - n2;
      v = 0.5 * (AAr[2 * (a * n1 + b) + t] - AAr[2 * (b * n1 + a) + t]);
    } else if (i < n2) {                                // BA[i, j - n2] = Xi[U2[i], U1[j - n2]]
      v = Xi[U1[2 * (j - n2)] * nx + U2[2 * i + t]];
    } else {                                            // -BA^T
      v = -Xi[U1[2 * (i - n2)] * nx + U2[2 * j + t]];
    }
    jb.N[idx] = v;
  }
}

inline size_t pf_site_smem() {
  return sizeof(double) * ((size_t)PFS_MAX * PFS_MAX + 3 * PFS_NX * PFS_NX + PFS_MAX + PFS_MAX + 2 + (PFS_MAX / 2) * 99 + 40) +
         sizeof(int) * (4 + 160 + 8);
}
}  // namespace tmf

extern "C" int tmf_pfaffian_site_finish(const tmf_pf_site_job *jobs_host, int njobs, void *desc_dev, void *stream) {
  using namespace tmf;
  if (njobs <= 0) return TMF_OK;
  int rc = copy_h2d(desc_dev, jobs_host, sizeof(tmf_pf_site_job) * (size_t)njobs, stream);
  if (rc) return rc;
  return launch_t("pf_site", pf_site_kernel, njobs, 256, pf_site_smem(), stream,
                  reinterpret_cast<const tmf_pf_site_job *>(desc_dev));
}
