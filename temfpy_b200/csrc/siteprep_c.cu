// K8 + K9 for complex Slater determinants: the nested-projector site stage of siteprep.cu in complex128, with the
// elimination of the always-occupied entangled orbitals folded into the same kernel.
//
// reference: slater.py:1071-1090 with complex v_bra / v_ket (O = HT(v_bra) @ v_ket, det + Schur complement).  The
// mode matrices come from tmf_slater_modes_nested_emb (interleaved complex columns), `a_col` is the complex column
// of C that couples the new site to the bra block (read from the embedded matrix, where row 2e holds its
// conjugate-transposed row: C[r, e] = conj(C[e, r])).  Same closed form as the real kernel,
//     S = Y_b^H Y_k - (Z_b^H Z_k)(1 - Z_k^H Z_k)^-1,   |det_always| = sqrt(det(1 - Z_k^H Z_k)),
// with Hermitian conjugates in place of transposes.  Not a throughput kernel (complex inputs are the rarer case):
// plain CTA-parallel loops, matrices of at most 34 x 34 in shared memory.
#include "cplx.hpp"

namespace tmf {

constexpr int CS_RC = 16;            // rows per staged chunk
constexpr int CS_MAX_MODES = 32;     // complex modes per bond (TMF_MAX_MODES real columns of the embedding)
constexpr double CS_MIN_PIVOT = 1e-9;

struct CFrame {
  int tr, R, Cc, ncand, k, s_r, s_c;
};
TMF_HD CFrame make_cframe(const tmf_site_job &jb) {
  CFrame f;
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

TMF_GLOBAL nested_site_c_kernel(const tmf_site_job *jobs, const tmf_nested_job *njobs) {
  const tmf_site_job jb = jobs[BLOCK_ID];
  const tmf_nested_job nj = njobs[BLOCK_ID];
  const CFrame fr = make_cframe(jb);
  const int kb = nj.k_bra, kk = nj.k_ket, ck = kk + nj.df, rb = kb + 1;
  const int n_b = jb.n_bra, n_k = n_b + 1;
  const int off = (jb.mode == 1) ? 1 : 0, edge = (jb.mode == 1) ? 0 : n_k - 1;
  const int ldq = ck + 1;
  const cplx *Vb = reinterpret_cast<const cplx *>(jb.Vb), *Vk = reinterpret_cast<const cplx *>(jb.Vk);
  const cplx *acol = reinterpret_cast<const cplx *>(nj.a_col);
  DYN_SMEM(unsigned char, raw);
  cplx *Q = reinterpret_cast<cplx *>(raw);          // (kb + 1) x (ck + 1)
  cplx *y0 = Q + (size_t)(kb + 1) * ldq;            // ck
  cplx *aZ = y0 + (ck + 1);                         // ck
  cplx *XZ = aZ + (ck + 1);                         // kb x ck
  cplx *YZ = XZ + (size_t)kb * ck + 1;              // ck x ck
  cplx *H = YZ + (size_t)ck * ck + 1;               // ck x ck
  cplx *K = H + (size_t)ck * ck + 1;                // rb x ck
  double *lamk = reinterpret_cast<double *>(K + (size_t)rb * ck + 1);   // ck
  double *eb = lamk + (ck + 1);                     // kb
  double *red = eb + (kb + 1);                      // 40
  int *ired = reinterpret_cast<int *>(red + 40);    // 40
  const int nbc = kb + 1, ncol = nbc + ck, lds = CS_RC + 1;
  cplx *st = XZ;                                    // staging overlays the phase-2 matrices

  PAR_FOR(idx, (kb + 1) * ldq) Q[idx] = cmake(0.0);
  CTA_SYNC();
  for (int r0 = 0; r0 < n_b; r0 += CS_RC) {
    PAR_FOR(idx, ncol * CS_RC) {
      const int col = idx / CS_RC, r = idx - col * CS_RC;
      const int gr = r0 + r;
      cplx v = cmake(0.0);
      if (gr < n_b) {
        if (col < kb) v = Vb[(int64_t)col * jb.ldb + gr];
        else if (col == kb) v = acol[gr];
        else v = Vk[(int64_t)(col - nbc) * jb.ldk + gr + off];
      }
      st[col * lds + r] = v;
    }
    CTA_SYNC();
    PAR_FOR(idx, (kb + 1) * (ck + 1)) {
      const int i = idx / (ck + 1), j = idx - i * (ck + 1);
      const cplx *bi = st + (size_t)i * lds, *kj = st + (size_t)(j < ck ? nbc + j : kb) * lds;
      cplx acc = cmake(0.0);
      for (int r = 0; r < CS_RC; ++r) acc = cadd(acc, cmulc(bi[r], kj[r]));
      Q[i * ldq + j] = cadd(Q[i * ldq + j], acc);
    }
    CTA_SYNC();
  }
  const int side = (jb.mode == 1) ? TMF_SIDE_R : TMF_SIDE_L;
  PAR_FOR(j, ck) {
    y0[j] = Vk[(int64_t)j * jb.ldk + edge];
    lamk[j] = (j < kk) ? ((side == TMF_SIDE_L) ? nj.e_ket[j] : 1.0 - nj.e_ket[j]) : 1.0;
  }
  PAR_FOR(i, kb) eb[i] = (side == TMF_SIDE_L) ? nj.e_bra[i] : 1.0 - nj.e_bra[i];
  CTA_SYNC();
  const double aa = Q[kb * ldq + ck].x;
  // X_b^H Z_k, a^H Z_k, Y'^H Z_k   (T = Q[:kb, :ck], w = Q[:kb, ck] = X_b^H a, ua = Q[kb, :ck] = a^H Y')
  PAR_FOR(idx, kb * ck) {
    const int i = idx / ck, j = idx - i * ck;
    XZ[idx] = csub(cscale(Q[i * ldq + j], lamk[j] - eb[i]), cmul(Q[i * ldq + ck], y0[j]));
  }
  PAR_FOR(j, ck) {
    cplx v = csub(cscale(Q[kb * ldq + j], lamk[j]), cscale(y0[j], aa));
    for (int i = 0; i < kb; ++i) v = csub(v, cscale(cmulc(Q[i * ldq + ck], Q[i * ldq + j]), eb[i]));
    aZ[j] = v;
  }
  PAR_FOR(idx, ck * ck) {
    const int j = idx / ck, l = idx - j * ck;
    cplx v = cscale(csub(cmake(j == l ? 1.0 : 0.0), cmulc(y0[j], y0[l])), lamk[l]);
    v = csub(v, cmulc(Q[kb * ldq + j], y0[l]));
    for (int i = 0; i < kb; ++i) v = csub(v, cscale(cmulc(Q[i * ldq + j], Q[i * ldq + l]), eb[i]));
    YZ[idx] = v;
  }
  CTA_SYNC();
  // G = Z_k^H Z_k,  K = Z_b^H Z_k
  PAR_FOR(idx, ck * ck) {
    const int j = idx / ck, l = idx - j * ck;
    cplx v = csub(cscale(YZ[idx], lamk[j]), cmulc(y0[j], aZ[l]));
    for (int i = 0; i < kb; ++i) v = csub(v, cscale(cmulc(Q[i * ldq + j], XZ[i * ck + l]), eb[i]));
    H[idx] = v;
  }
  PAR_FOR(idx, rb * ck) {
    const int r = idx / ck, l = idx - r * ck;
    cplx v;
    if (r == 0) {
      v = aZ[l];
      for (int j = 0; j < ck; ++j) v = csub(v, cscale(cmul(y0[j], YZ[j * ck + l]), lamk[j]));
    } else {
      const int i = r - 1;
      v = cscale(XZ[i * ck + l], eb[i]);
      for (int j = 0; j < ck; ++j) v = csub(v, cscale(cmul(Q[i * ldq + j], YZ[j * ck + l]), lamk[j]));
    }
    K[idx] = v;
  }
  CTA_SYNC();
  // YZ <- H = 1 - (G + G^H) / 2, Cholesky H = L L^H (lower triangle, row-major)
  PAR_FOR(idx, ck * ck) {
    const int j = idx / ck, l = idx - j * ck;
    const cplx g = cscale(cadd(H[j * ck + l], cconj(H[l * ck + j])), 0.5);
    YZ[idx] = csub(cmake(j == l ? 1.0 : 0.0), g);
  }
  PAR_FOR(one, 1) { red[0] = 1.0; red[1] = 0.0; red[2] = 1.0; red[3] = 0.0; red[4] = 1.0; }
  CTA_SYNC();
  cplx *Lc = YZ;
  for (int j = 0; j < ck; ++j) {
    PAR_FOR(one, 1) {
      const double d = Lc[j * ck + j].x;
      if (!(d > 1e-3)) red[1] = 1.0;
      const double sd = sqrt(d > 1e-300 ? d : 1e-300);
      Lc[j * ck + j] = cmake(sd);
      red[0] *= sd;
    }
    CTA_SYNC();
    const double inv = 1.0 / Lc[j * ck + j].x;
    PAR_FOR(i, ck - j - 1) Lc[(j + 1 + i) * ck + j] = cscale(Lc[(j + 1 + i) * ck + j], inv);
    CTA_SYNC();
    const int m = ck - j - 1;
    PAR_FOR(idx, m * m) {
      const int i = j + 1 + idx / m, l = j + 1 + idx % m;
      if (l <= i) Lc[i * ck + l] = csub(Lc[i * ck + l], cmul(Lc[i * ck + j], cconj(Lc[l * ck + j])));
    }
    CTA_SYNC();
  }
  // rows of K <- S0 - K H^-1   (x L L^H = k: forward with L^H, backward with L)
  PAR_FOR(r, rb) {
    cplx *x = K + (size_t)r * ck;
    for (int j = 0; j < ck; ++j) {
      cplx v = x[j];
      for (int l = 0; l < j; ++l) v = csub(v, cmul(x[l], cconj(Lc[j * ck + l])));
      x[j] = cscale(v, 1.0 / Lc[j * ck + j].x);
    }
    for (int j = ck - 1; j >= 0; --j) {
      cplx v = x[j];
      for (int l = j + 1; l < ck; ++l) v = csub(v, cmul(x[l], Lc[l * ck + j]));
      x[j] = cscale(v, 1.0 / Lc[j * ck + j].x);
    }
    for (int j = 0; j < ck; ++j) x[j] = csub((r == 0) ? y0[j] : Q[(r - 1) * ldq + j], x[j]);
  }
  CTA_SYNC();
  // ---- frame of the always-orbital elimination (orbital order and signs of the site plan) -----------
  cplx *O = reinterpret_cast<cplx *>(jb.O);
  const int nbo = jb.ka_bra + jb.sb, nko = jb.ka_ket + jb.sk, ld = fr.R;
  PAR_FOR(idx, nbo * nko) {
    const int pb = idx % nbo, n = idx / nbo;
    const int cb = jb.bra_cols[pb], cn = jb.ket_cols[n];
    const int rbn = (cb < 0) ? 0 : 1 + cb;
    // frame entries are <bra orbital | ket orbital>; when the ket side provides the rows the entry is transposed
    // (not conjugated: the determinant of the transposed overlap matrix is the same number)
    const cplx v = cscale(K[(size_t)rbn * ck + cn], jb.bra_sign[pb] * jb.ket_sign[n]);
    if (fr.tr) O[(int64_t)pb * ld + n] = v;
    else O[(int64_t)n * ld + pb] = v;
  }
  CTA_SYNC();
  // ---- complex LU with row pivoting among the always orbitals of the row side (slater.py:1077-1090) ---
  const int rows = fr.R, cols = fr.Cc, kel = fr.k;
  for (int t = 0; t < kel; ++t) {
    PAR_FOR(lane, 32) {
      double best = -1.0;
      int bi = t;
      for (int r = t + lane; r < fr.ncand; r += 32) {
        const double a = cabs2(O[(int64_t)t * ld + r]);
        if (a > best) { best = a; bi = r; }
      }
      red[8 + lane] = best;
      ired[lane] = bi;
    }
    CTA_SYNC();
    PAR_FOR(one, 1) {
      double best = red[8];
      int bi = ired[0];
      for (int l = 1; l < 32; ++l)
        if (red[8 + l] > best || (red[8 + l] == best && ired[l] < bi)) { best = red[8 + l]; bi = ired[l]; }
      ired[32] = bi;
      const cplx pv = O[(int64_t)t * ld + bi];
      cplx d = cmul(cmake(red[2], red[3]), pv);
      if (bi != t) d = cneg(d);
      red[2] = d.x; red[3] = d.y;
      const double ap = sqrt(cabs2(pv));
      red[4] = ap < red[4] ? ap : red[4];
      const cplx iv = cinv(pv);
      red[5] = iv.x; red[6] = iv.y;
    }
    CTA_SYNC();
    const int p = ired[32];
    if (p != t) {
      PAR_FOR(c, cols) {
        const cplx a = O[(int64_t)c * ld + t];
        O[(int64_t)c * ld + t] = O[(int64_t)c * ld + p];
        O[(int64_t)c * ld + p] = a;
      }
      CTA_SYNC();
    }
    const cplx iv = cmake(red[5], red[6]);
    PAR_FOR(r, rows - t - 1) O[(int64_t)t * ld + t + 1 + r] = cmul(O[(int64_t)t * ld + t + 1 + r], iv);
    CTA_SYNC();
    const int hr = rows - t - 1, hc = cols - t - 1;
    PAR_FOR(idx, hr * hc) {
      const int c = t + 1 + idx / hr, r = t + 1 + idx % hr;
      O[(int64_t)c * ld + r] = csub(O[(int64_t)c * ld + r], cmul(O[(int64_t)t * ld + r], O[(int64_t)c * ld + t]));
    }
    CTA_SYNC();
  }
  // ---- Schur complement -> S in the reference's row / column order (as schur_kernel) ------------------
  cplx *S = reinterpret_cast<cplx *>(jb.S);
  const int nlo = fr.ncand - kel, nfr = nlo + fr.s_r, nfc = fr.s_c;
  const int s_bra = fr.tr ? nfc : nfr;
  PAR_FOR(idx, nfr * nfc) {
    const int c = idx / nfr, r = idx - c * nfr;
    const int ref_r = (jb.mode == 0) ? r : (r < nlo ? fr.s_r + r : r - nlo);
    const cplx v = O[(int64_t)(kel + c) * ld + kel + r];
    if (fr.tr) S[(int64_t)ref_r * s_bra + c] = v;
    else S[(int64_t)c * s_bra + ref_r] = v;
  }
  PAR_FOR(one, 1) {
    cplx d = cscale(cmake(red[2], red[3]), red[0]);
    if (red[1] != 0.0 || red[4] < CS_MIN_PIVOT) d = cmake(nan(""), nan(""));
    jb.det[0] = d.x;
    jb.det[1] = d.y;
  }
}

static size_t nested_c_smem_bytes(int kb, int ck) {
  const size_t head = (size_t)(kb + 1) * (ck + 1) + 2 * (size_t)(ck + 1);
  const size_t phase2 = (size_t)kb * ck + 2 * (size_t)ck * ck + (size_t)(kb + 1) * ck + 4;
  const size_t stage = (size_t)(kb + 1 + ck) * (CS_RC + 1);
  return sizeof(cplx) * (head + std::max(phase2, stage) + 4) + sizeof(double) * ((ck + 1) + (kb + 1) + 40) +
         sizeof(int) * 48 + 64;
}

}  // namespace tmf

// Complex form of tmf_site_nested_batched: V slots hold interleaved complex columns (ldb / ldk in complex
// elements), O / S are complex (2 doubles per entry), det is complex (2 doubles per site).
extern "C" int tmf_site_nested_c_batched(const tmf_site_job *jobs_host, const tmf_nested_job *njobs_host,
                                         int nsites, void *desc_dev, void *stream) {
  using namespace tmf;
  if (nsites <= 0) return TMF_OK;
  size_t smem = 0;
  for (int s = 0; s < nsites; ++s) {
    const tmf_nested_job &nj = njobs_host[s];
    if (nj.df < 0 || nj.df > 1 || nj.k_bra < 0 || nj.k_ket < 0 || nj.k_bra > CS_MAX_MODES || nj.k_ket > CS_MAX_MODES ||
        !jobs_host[s].physical) {
      set_error("tmf_site_nested_c_batched: more than 32 complex modes per bond, or filled spaces not nested");
      return TMF_ERR_VALUE;
    }
    smem = std::max(smem, nested_c_smem_bytes(nj.k_bra, nj.k_ket + nj.df));
  }
  unsigned char *d = static_cast<unsigned char *>(desc_dev);
  const size_t o_nest = align256(sizeof(tmf_site_job) * (size_t)nsites);
  int rc = copy_h2d(d, jobs_host, sizeof(tmf_site_job) * (size_t)nsites, stream);
  if (rc) return rc;
  rc = copy_h2d(d + o_nest, njobs_host, sizeof(tmf_nested_job) * (size_t)nsites, stream);
  if (rc) return rc;
  return launch_t("nested_site_c", nested_site_c_kernel, nsites, 256, smem, stream,
                  reinterpret_cast<const tmf_site_job *>(d), reinterpret_cast<const tmf_nested_job *>(d + o_nest));
}
