// Canonical form of a finite, charge-conserving MPS on the device -- the job of TeNPy's
// `MPS.canonical_form_finite` as called after the Gutzwiller projection (reference gutzwiller.py:262-267 / :467-472):
// a left-to-right QR sweep followed by a right-to-left SVD sweep, block by block in the charge sectors.
//
// The sweep is sequential in the sites, so the work of one step is a handful of small factorisations (one per charge
// sector of a bond) plus the product that carries the triangular / singular factor into the neighbouring tensor.  A
// step is three launches on one stream, with no host synchronisation anywhere in the sweep (all block sizes follow
// from the sector tables, so the whole schedule is planned on the host beforehand; the singular values below `cutoff`
// are removed by the caller at the very end, which is equivalent to the reference's truncation on the fly because a
// discarded direction only ever multiplies zeros afterwards):
//   1. block_qr_kernel / block_svd_kernel: one CTA per charge sector.  The block is gathered from its two panels
//      (the two physical values) into a workspace, orthogonalised by Gram-Schmidt with re-orthogonalisation (rank
//      deficient blocks are the rule after a projection: dependent columns become zero columns) and, in the SVD
//      sweep, diagonalised by the one-sided Jacobi routine of the mode extraction on the small triangular factor;
//   2. inv_norm_kernel: the Frobenius norm that couples the sectors of a bond (one scalar; the next step's kernels
//      scale their input by it while loading);
//   3. the grouped DMMA GEMM: R (or U S) times the blocks of the neighbouring tensor, written into the tensor's
//      buffer for the new bond dimension.
// Tensors are dense T[vL, p, vR] (row-major), sectors are contiguous index ranges of a bond, q(vL) + qp[p] = q(vR).
#include "jacobi.cuh"

#include <algorithm>
#include <map>

namespace tmf {
int gemm_grouped(const tmf_gemm_job *jobs, int njobs, void *desc_dev, void *stream);
int64_t gemm_desc_bytes(int njobs);

constexpr int CANON_MAX = 160;        // largest min(rows, cols) of a block (Jacobi scratch in shared memory)

struct BlockQrJob {
  const double *src[2];   // panel t (m[t] rows): element (r, c) at src[t][r * rs + c]
  double *dst[2];         // Q, panel t: element (r, i) at dst[t][r * rd + i], i < k
  double *R;              // k x n column-major (ld = k)
  double *work;           // (m0 + m1) * n doubles
  double *nrm2;           // ||R||_F^2 of this block
  const double *scale;    // optional device scalar multiplied into the input
  int64_t rs, rd;
  int m[2], n, k;
  int pad_[8];
};
struct BlockSvdJob {
  const double *src[2];   // panel t (n[t] columns): element (r, c) at src[t][r * rs + c], r < m
  double *dst[2];         // Vh, panel t: element (i, c) at dst[t][i * rd + c], i < k
  double *US;             // m x k column-major (ld = m): U diag(S)
  double *S;              // k singular values, decreasing
  double *work;           // svd_work_doubles(m, n0 + n1)
  double *nrm2;           // sum of S^2 of this block
  const double *scale;
  int64_t rs, rd;
  int m, n[2], k;
  const double *csq;      // optional: column c of the input is multiplied by csq[c]^2 (Procrustes weights)
  int want_u, pad_[3];    // want_u: the US buffer receives U instead of U diag(S)
};
static_assert(sizeof(BlockQrJob) == 128 && sizeof(BlockSvdJob) == 128, "descriptors must be 128 bytes");

// Gram-Schmidt QR with re-orthogonalisation of A (ra x ca, column-major, ld = ra) by one CTA.  On exit the first
// `live` columns of A are orthonormal, the others zero, and A_in = A_out * R with R (kmax x ca, column-major,
// ld = kmax, zeroed by the caller).  A column whose remainder after two projections is below 1e-13 of its norm is
// dependent: it contributes coefficients to R but no direction.  part: 33 * max(ca, 1) doubles, coef: ca doubles.
TMF_DEVICE int cta_gs_qr(double *A, int ra, int ca, int kmax, double *R, double *part, double *coef) {
  int live = 0;
  for (int c = 0; c < ca; ++c) {
    double *v = A + (int64_t)c * ra;
    double n0 = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
      // item `live` is v . v (norm before the pass), items i < live are q_i . v
      PAR_FOR(item, (live + 1) * 32) {
        const int i = item >> 5, lane = item & 31;
        const double *q = (i < live) ? A + (int64_t)i * ra : v;
        double s = 0.0;
        for (int r = lane; r < ra; r += 32) s += q[r] * v[r];
        part[i * 33 + lane] = s;
      }
      CTA_SYNC();
      PAR_FOR(i, live + 1) {
        double s = 0.0;
        for (int l = 0; l < 32; ++l) s += part[i * 33 + l];
        coef[i] = s;
        if (i < live) R[i + (int64_t)kmax * c] += s;
      }
      CTA_SYNC();
      if (pass == 0) n0 = coef[live];
      if (live > 0) {
        PAR_FOR(r, ra) {
          double s = v[r];
          for (int i = 0; i < live; ++i) s -= A[(int64_t)i * ra + r] * coef[i];
          v[r] = s;
        }
      }
      CTA_SYNC();
      if (live == 0) break;
    }
    PAR_FOR(lane, 32) {
      double s = 0.0;
      for (int r = lane; r < ra; r += 32) s += v[r] * v[r];
      part[lane] = s;
    }
    CTA_SYNC();
    double n1 = 0.0;
    for (int l = 0; l < 32; ++l) n1 += part[l];
    const bool ok = live < kmax && n1 > 1e-26 * n0 && n1 > 0.0;
    const double inv = ok ? 1.0 / sqrt(n1) : 0.0;
    double *q = A + (int64_t)live * ra;
    CTA_SYNC();
    if (ok) {
      PAR_FOR(r, ra) {
        const double x = v[r] * inv;
        if (live != c) v[r] = 0.0;
        q[r] = x;
      }
      PAR_FOR(one, 1) R[live + (int64_t)kmax * c] = sqrt(n1);
      ++live;
    } else {
      PAR_FOR(r, ra) v[r] = 0.0;
    }
    CTA_SYNC();
  }
  return live;
}

// nmax: largest n of the launch (sizes the scratch), mat_doubles: shared memory left for the matrices -- a block whose
// m x n copy and k x n factor fit works entirely in shared memory (every phase of the Gram-Schmidt loop is a
// round trip to wherever the matrix lives), larger ones in their global workspace.
TMF_GLOBAL block_qr_kernel(const BlockQrJob *jobs, int nmax, int mat_doubles) {
  const BlockQrJob jb = jobs[BLOCK_ID];
  const int m0 = jb.m[0], m = jb.m[0] + jb.m[1], n = jb.n, k = jb.k;
  DYN_SMEM(double, sm);
  double *part = sm;                          // 33 * (nmax + 1)
  double *coef = part + 33 * (nmax + 1);      // nmax + 40
  double *mat = coef + nmax + 40;
  const bool in_smem = (int64_t)m * n + (int64_t)k * n <= mat_doubles;
  double *W = in_smem ? mat : jb.work;
  double *Rw = in_smem ? mat + (int64_t)m * n : jb.R;
  const double sc = jb.scale ? *jb.scale : 1.0;
  PAR_FOR(idx, m * n) {
    const int r = idx / n, c = idx - r * n;
    const int t = r >= m0, rr = t ? r - m0 : r;
    W[(int64_t)c * m + r] = sc * jb.src[t][rr * jb.rs + c];
  }
  PAR_FOR(idx, k * n) Rw[idx] = 0.0;
  CTA_SYNC();
  cta_gs_qr(W, m, n, k, Rw, part, coef);
  PAR_FOR(idx, m * k) {
    const int r = idx / k, i = idx - r * k;
    const int t = r >= m0, rr = t ? r - m0 : r;
    jb.dst[t][rr * jb.rd + i] = W[(int64_t)i * m + r];
  }
  if (in_smem) {
    PAR_FOR(idx, k * n) jb.R[idx] = Rw[idx];
    CTA_SYNC();
  }
  PAR_FOR(lane, 32) {
    double s = 0.0;
    for (int i = lane; i < k * n; i += 32) s += jb.R[i] * jb.R[i];
    coef[lane] = s;
  }
  CTA_SYNC();
  PAR_FOR(one, 1) {
    double s = 0.0;
    for (int l = 0; l < 32; ++l) s += coef[l];
    *jb.nrm2 = s;
  }
}
constexpr int64_t CANON_SMEM_BYTES = 200 * 1024;
inline size_t block_qr_scratch(int nmax) { return sizeof(double) * ((size_t)33 * (nmax + 1) + nmax + 40); }

inline int64_t svd_work_doubles(int m, int n) {
  const int64_t ra = std::max(m, n), ca = std::min(m, n);
  return ra * ca + 2 * ca * ca + 64;
}

// camax: largest min(m, n) of the launch (sizes the scratch); mat_doubles: shared memory left for A, R1 and J
TMF_GLOBAL block_svd_kernel(const BlockSvdJob *jobs, int camax, int mat_doubles) {
  const BlockSvdJob jb = jobs[BLOCK_ID];
  const int m = jb.m, n0 = jb.n[0], n = jb.n[0] + jb.n[1], k = jb.k;
  const bool tall = m >= n;                       // A = M (m x n) or A = M^T (n x m): ra >= ca = k
  const int ra = tall ? m : n, ca = tall ? n : m;
  DYN_SMEM(double, sm);
  const int npm = (camax + 1) & ~1;
  double *rot = sm;                                // npm
  double *part = rot + npm + 2;                    // max(33 * (camax + 1), (npm / 2) * 99)
  const int npa = 33 * (camax + 1), npb = (npm / 2) * 99;
  const int npart = (npa > npb ? npa : npb) + 8;
  double *coef = part + npart;                     // camax + 40
  double *sig = coef + camax + 40;                 // camax
  int *iw = reinterpret_cast<int *>(sig + camax + 2);
  int *flag = iw, *sel = iw + 2, *rank = sel + camax, *cnt = rank + camax;
  double *mat = reinterpret_cast<double *>(iw + ((3 * camax + 34) & ~1));
  const bool in_smem = (int64_t)ra * ca + 2 * (int64_t)ca * ca <= mat_doubles;
  double *A = in_smem ? mat : jb.work, *R1 = A + (int64_t)ra * ca, *J = R1 + (int64_t)ca * ca;
  const double sc = jb.scale ? *jb.scale : 1.0;
  // columns of A in order of decreasing norm (de Rijk): with the Jacobi sweeps on the transposed triangular factor
  // below this is what makes a graded block converge in a handful of sweeps
  int *cperm = cnt + 8;                            // camax
  PAR_FOR(item, ca * 32) {
    const int j = item >> 5, lane = item & 31;
    double s2 = 0.0;
    if (tall) {
      const int t = j >= n0, cc = t ? j - n0 : j;
      const double w = jb.csq ? jb.csq[j] * jb.csq[j] : 1.0;
      for (int r = lane; r < m; r += 32) { const double v = w * jb.src[t][r * jb.rs + cc]; s2 += v * v; }
    } else {
      for (int c = lane; c < n; c += 32) {
        const int t = c >= n0, cc = t ? c - n0 : c;
        const double v = (jb.csq ? jb.csq[c] * jb.csq[c] : 1.0) * jb.src[t][j * jb.rs + cc];
        s2 += v * v;
      }
    }
    part[j * 33 + lane] = s2;
  }
  CTA_SYNC();
  PAR_FOR(j, ca) {
    double s2 = 0.0;
    for (int l = 0; l < 32; ++l) s2 += part[j * 33 + l];
    sig[j] = s2;
    sel[j] = 1;
  }
  CTA_SYNC();
  rank_desc(sig, sel, ca, cperm, cnt);
  PAR_FOR(idx, m * n) {
    const int r = idx / n, c = idx - r * n;
    const int t = c >= n0, cc = t ? c - n0 : c;
    const double v = sc * (jb.csq ? jb.csq[c] * jb.csq[c] : 1.0) * jb.src[t][r * jb.rs + cc];
    if (tall) A[(int64_t)cperm[c] * ra + r] = v; else A[(int64_t)cperm[r] * ra + c] = v;
  }
  PAR_FOR(idx, ca * ca) {
    R1[idx] = 0.0;
    J[idx] = (idx / ca == idx % ca) ? 1.0 : 0.0;
  }
  CTA_SYNC();
  cta_gs_qr(A, ra, ca, ca, R1, part, coef);        // A = Q1 (ra x ca), A_in = Q1 R1
  // The Jacobi routine treats columns with squared norm < 1e-30 as null (rotations with them would never settle);
  // that test is absolute, so the factor is brought to unit Frobenius norm first.  The blocks of a projected state
  // routinely carry weights many orders of magnitude apart.
  PAR_FOR(lane, 32) {
    double s = 0.0;
    for (int i = lane; i < ca * ca; i += 32) s += R1[i] * R1[i];
    part[lane] = s;
  }
  CTA_SYNC();
  double fro = 0.0;
  for (int l = 0; l < 32; ++l) fro += part[l];
  fro = sqrt(fro);
  CTA_SYNC();
  if (fro > 0.0) {
    // Jacobi on the *transposed* factor (Drmac / Veselic): the rows of the triangular factor of a graded matrix are
    // graded themselves, and the one-sided sweeps on L = R1^T settle in a few passes where those on R1 need dozens.
    const double rf = 1.0 / fro;
    PAR_FOR(idx, ca * ca) {
      const int c = idx / ca, r = idx - c * ca;
      if (r < c) {
        const double a = R1[(int64_t)c * ca + r] * rf, b = R1[(int64_t)r * ca + c] * rf;
        R1[(int64_t)c * ca + r] = b;
        R1[(int64_t)r * ca + c] = a;
      } else if (r == c) {
        R1[idx] *= rf;
      }
    }
    CTA_SYNC();
    jacobi_onesided(R1, ca, J, ca, ca, rot, part, flag);   // G = R1^T:  G J = W diag(sigma)
    PAR_FOR(idx, ca * ca) R1[idx] *= fro;
    CTA_SYNC();
  }
  PAR_FOR(i, ca) {
    double s = 0.0;
    for (int r = 0; r < ca; ++r) s += R1[(int64_t)i * ca + r] * R1[(int64_t)i * ca + r];
    sig[i] = sqrt(s);
    sel[i] = 1;
  }
  CTA_SYNC();
  rank_desc(sig, sel, ca, rank, cnt);
  // R1 = J diag(sigma) W^T, so A_in = (Q1 J) diag(sigma) W^T with W diag(sigma) in the R1 buffer (A_in: columns permuted).
  // tall: M = A_in:   U = Q1 J, Vh = W^T;      wide: M = A_in^T:   U = W, Vh = (Q1 J)^T
  if (tall) {
    PAR_FOR(idx, m * k) {
      const int i = idx / m, r = idx - i * m;
      double s = 0.0;
      for (int t = 0; t < ca; ++t) s += A[(int64_t)t * ra + r] * J[(int64_t)i * ca + t];
      jb.US[(int64_t)rank[i] * m + r] = jb.want_u ? s : s * sig[i];
    }
    PAR_FOR(idx, k * n) {
      const int i = idx / n, c = idx - i * n;
      const int t = c >= n0, cc = t ? c - n0 : c;
      jb.dst[t][rank[i] * jb.rd + cc] = (sig[i] > 0.0) ? R1[(int64_t)i * ca + cperm[c]] / sig[i] : 0.0;
    }
  } else {
    PAR_FOR(idx, m * k) {
      const int i = idx / m, r = idx - i * m;
      const double w = R1[(int64_t)i * ca + cperm[r]];
      jb.US[(int64_t)rank[i] * m + r] = jb.want_u ? (sig[i] > 0.0 ? w / sig[i] : 0.0) : w;
    }
    PAR_FOR(idx, k * n) {
      const int i = idx / n, c = idx - i * n;
      const int t = c >= n0, cc = t ? c - n0 : c;
      double s = 0.0;
      for (int u = 0; u < ca; ++u) s += A[(int64_t)u * ra + c] * J[(int64_t)i * ca + u];
      jb.dst[t][rank[i] * jb.rd + cc] = s;
    }
  }
  PAR_FOR(i, k) jb.S[rank[i]] = sig[i];
  PAR_FOR(one, 1) {
    double s = 0.0;
    for (int i = 0; i < k; ++i) s += sig[i] * sig[i];
    *jb.nrm2 = s;
  }
}
inline size_t block_svd_scratch(int camax) {
  const int np = (camax + 1) & ~1;
  const size_t npart = std::max(33 * (camax + 1), (np / 2) * 99) + 8;
  return sizeof(double) * ((size_t)np + 2 + npart + camax + 40 + camax + 2) + sizeof(int) * (size_t)((3 * camax + 34) & ~1);
}

// Orthogonal Procrustes per charge sector (reference iMPS.py:150-184, K15): R = U Vh of M = C diag(sk^2) from the
// SVD kernel above (want_u), written into the dense rotation matrix, together with the two sums behind the error
// metrics of iMPS.py:139-147 / :186-190:  metrics[0] = sum |C sk|^2,  metrics[1] = sum |(R - C) sk|^2.
struct ProcrustesFinish {
  const double *C, *sk, *U, *Vh, *S;
  double *R, *metrics;
  int64_t ldc, ldr;
  int m, n, k, pad_;
};
TMF_GLOBAL procrustes_finish_kernel(const ProcrustesFinish *jobs) {
  const ProcrustesFinish jb = jobs[BLOCK_ID];
  DYN_SMEM(double, part);       // 2 * NTHREADS
  PAR_FOR(t, NTHREADS) {
    double a = 0.0, b = 0.0;
    for (int idx = t; idx < jb.m * jb.n; idx += NTHREADS) {
      const int r = idx / jb.n, c = idx - r * jb.n;
      double s = 0.0;
      for (int i = 0; i < jb.k; ++i)
        if (jb.S[i] > 0.0) s += jb.U[(int64_t)i * jb.m + r] * jb.Vh[(int64_t)i * jb.n + c];
      jb.R[r * jb.ldr + c] = s;
      const double cv = jb.C[r * jb.ldc + c], w = jb.sk[c];
      a += cv * w * cv * w;
      b += (s - cv) * w * (s - cv) * w;
    }
    part[t] = a;
    part[NTHREADS + t] = b;
  }
  CTA_SYNC();
  PAR_FOR(one, 1) {
    double a = 0.0, b = 0.0;
    for (int t = 0; t < NTHREADS; ++t) { a += part[t]; b += part[NTHREADS + t]; }
    jb.metrics[0] = a;
    jb.metrics[1] = b;
  }
}

// out[0] = 1 / sqrt(sum) with sum = sum_i vals[i] (squares == 0) or sum_i vals[i]^2 (squares == 1); 1 if the sum is 0
TMF_GLOBAL inv_norm_kernel(const double *vals, int64_t n, int squares, double *out) {
  DYN_SMEM(double, part);
  PAR_FOR(t, NTHREADS) {
    double s = 0.0;
    for (int64_t i = t; i < n; i += NTHREADS) s += squares ? vals[i] * vals[i] : vals[i];
    part[t] = s;
  }
  CTA_SYNC();
  PAR_FOR(one, 1) {
    double s = 0.0;
    for (int t = 0; t < NTHREADS; ++t) s += part[t];
    out[0] = (s > 0.0) ? 1.0 / sqrt(s) : 1.0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: schedule of the two sweeps
// ---------------------------------------------------------------------------------------------------------------
struct Sector { int q, start, count; };
typedef std::vector<Sector> Sectors;

static bool sectors_of(const int *q, int n, Sectors &out) {
  out.clear();
  for (int i = 0; i < n;) {
    int j = i;
    while (j < n && q[j] == q[i]) ++j;
    if (!out.empty() && q[i] <= out.back().q) return false;      // not sorted: sectors would not be contiguous
    out.push_back({q[i], i, j - i});
    i = j;
  }
  return true;
}
static const Sector *find_sector(const Sectors &s, int q) {
  for (const Sector &x : s)
    if (x.q == q) return &x;
  return nullptr;
}
static int total(const Sectors &s) {
  int t = 0;
  for (const Sector &x : s) t += x.count;
  return t;
}
}  // namespace tmf

struct tmf_canon {
  int L = 0, qp[2] = {0, 0};
  std::vector<tmf::Sectors> s0, s1, s2;     // sectors of every bond: input, after the QR sweep, final
  // (old sector index feeding every new sector, needed to address the R / US blocks)
  std::vector<std::vector<int>> src1, src2;
  int cmax = 0;
  int64_t work_bytes = 0, t2_doubles = 0, s_doubles = 0;
};

namespace tmf {
struct CanonBuffers {
  std::vector<double *> Tcur, T1, Xcur;
  std::vector<std::vector<double *>> Rb, USb;
  double *scratch = nullptr, *nrm = nullptr, *inv1 = nullptr, *invx = nullptr;
  unsigned char *jobs = nullptr;
  void *gdesc = nullptr;
  int64_t njobs = 0, zero_bytes = 0, used = 0;
  bool ok = true;
};
// rows of the QR block of new sector u of bond j+1 / columns of the SVD block of new sector u of bond j
static int qr_rows(const tmf_canon *c, int j, size_t u) {
  const Sector &col = c->s0[j + 1][c->src1[j + 1][u]];
  int m = 0;
  for (int p = 0; p < 2; ++p)
    if (const Sector *r = find_sector(c->s1[j], col.q - c->qp[p])) m += r->count;
  return m;
}
static int svd_cols(const tmf_canon *c, int j, size_t u) {
  const Sector &row = c->s1[j][c->src2[j][u]];
  int n = 0;
  for (int p = 0; p < 2; ++p)
    if (const Sector *s = find_sector(c->s2[j + 1], row.q + c->qp[p])) n += s->count;
  return n;
}
// carves the workspace (base == nullptr: dry run for the size)
static void canon_layout(const tmf_canon *c, void *base, int64_t bytes, CanonBuffers &b) {
  const int L = c->L;
  Arena ar(base, bytes);
  b.Tcur.resize(L); b.T1.resize(L); b.Xcur.resize(L);
  b.Rb.assign(L + 1, {}); b.USb.assign(L + 1, {});
  for (int j = 0; j < L; ++j) {
    const int64_t d1 = total(c->s1[j]);
    b.Tcur[j] = ar.take<double>(d1 * 2 * total(c->s0[j + 1]));
    b.T1[j] = ar.take<double>(d1 * 2 * total(c->s1[j + 1]));
    b.Xcur[j] = ar.take<double>(d1 * 2 * total(c->s2[j + 1]));
  }
  int64_t sq_max = 0, gemm_max = 1;
  b.njobs = 0;
  for (int j = 0; j + 1 < L; ++j) {
    int64_t sq = 0;
    for (size_t u = 0; u < c->s1[j + 1].size(); ++u) {
      const int n = c->s0[j + 1][c->src1[j + 1][u]].count;
      b.Rb[j + 1].push_back(ar.take<double>((int64_t)c->s1[j + 1][u].count * n));
      sq += align256(8 * (int64_t)qr_rows(c, j, u) * n) / 8;
    }
    sq_max = std::max(sq_max, sq);
    b.njobs += (int64_t)c->s1[j + 1].size();
    gemm_max = std::max<int64_t>(gemm_max, 2 * (int64_t)c->s1[j + 1].size());
  }
  for (int j = L - 1; j >= 0; --j) {
    int64_t sq = 0;
    for (size_t u = 0; u < c->s2[j].size(); ++u) {
      const int m = c->s1[j][c->src2[j][u]].count;
      b.USb[j].push_back(ar.take<double>((int64_t)m * c->s2[j][u].count));
      sq += align256(8 * svd_work_doubles(m, svd_cols(c, j, u))) / 8;
    }
    sq_max = std::max(sq_max, sq);
    b.njobs += (int64_t)c->s2[j].size();
    gemm_max = std::max<int64_t>(gemm_max, 2 * (int64_t)c->s2[j].size());
  }
  b.scratch = ar.take<double>(sq_max + 8);
  b.zero_bytes = (int64_t)(reinterpret_cast<unsigned char *>(b.scratch) - static_cast<unsigned char *>(base));
  b.nrm = ar.take<double>(b.njobs + 8);
  b.inv1 = ar.take<double>(L + 2);
  b.invx = ar.take<double>(2);
  b.jobs = ar.take<unsigned char>(128 * (b.njobs + 1));
  b.gdesc = ar.take<unsigned char>(gemm_desc_bytes((int)gemm_max));
  b.used = ar.used;
  b.ok = ar.ok();
}
}  // namespace tmf

extern "C" tmf_canon *tmf_canon_create(int L, const int *dims0, const int *charges0, const int *qp) {
  using namespace tmf;
  if (L <= 0) { set_error("tmf_canon_create: empty MPS"); return nullptr; }
  tmf_canon *c = new tmf_canon();
  c->L = L; c->qp[0] = qp[0]; c->qp[1] = qp[1];
  c->s0.resize(L + 1); c->s1.resize(L + 1); c->s2.resize(L + 1);
  c->src1.resize(L + 1); c->src2.resize(L + 1);
  int64_t off = 0;
  for (int j = 0; j <= L; ++j) {
    if (!sectors_of(charges0 + off, dims0[j], c->s0[j])) {
      set_error("tmf_canon_create: the charges of a bond must be sorted (contiguous sectors)");
      delete c;
      return nullptr;
    }
    off += dims0[j];
  }
  auto fail = [&](const char *msg) { set_error(msg); delete c; return (tmf_canon *)nullptr; };
  // QR sweep (bonds 1 .. L-1 get a new basis; gutzwiller.py:266 -> canonical_form_finite)
  c->s1[0] = c->s0[0];
  for (int j = 0; j + 1 < L; ++j) {
    int pos = 0;
    for (size_t u = 0; u < c->s0[j + 1].size(); ++u) {
      const Sector &col = c->s0[j + 1][u];
      int m = 0;
      for (int p = 0; p < 2; ++p)
        if (const Sector *r = find_sector(c->s1[j], col.q - qp[p])) m += r->count;
      if (m == 0) continue;
      const int k = std::min(m, col.count);
      if (col.count > 640) return fail("tmf_canon_create: charge sector larger than 640 (host sweep required)");
      c->s1[j + 1].push_back({col.q, pos, k});
      c->src1[j + 1].push_back((int)u);
      pos += k;
    }
  }
  c->s1[L] = c->s0[L];
  // SVD sweep (bonds L-1 .. 0)
  c->s2[L] = c->s1[L];
  for (int j = L - 1; j >= 0; --j) {
    int pos = 0;
    for (size_t u = 0; u < c->s1[j].size(); ++u) {
      const Sector &row = c->s1[j][u];
      int n = 0;
      for (int p = 0; p < 2; ++p)
        if (const Sector *s = find_sector(c->s2[j + 1], row.q + qp[p])) n += s->count;
      if (n == 0) continue;
      const int k = std::min(row.count, n);
      if (k > CANON_MAX) return fail("tmf_canon_create: charge block larger than 160 (host sweep required)");
      c->cmax = std::max(c->cmax, k);
      c->s2[j].push_back({row.q, pos, k});
      c->src2[j].push_back((int)u);
      pos += k;
    }
  }
  for (int j = 0; j < L; ++j) {
    c->t2_doubles += (int64_t)total(c->s2[j]) * 2 * total(c->s2[j + 1]);
    c->s_doubles += total(c->s2[j]);
  }
  CanonBuffers dry;
  canon_layout(c, nullptr, 0, dry);
  c->work_bytes = dry.used + 4096;
  return c;
}

extern "C" void tmf_canon_destroy(tmf_canon *c) { delete c; }

extern "C" int tmf_canon_sizes(const tmf_canon *c, int64_t *q) {
  q[0] = c->work_bytes;
  q[1] = c->t2_doubles;
  q[2] = c->s_doubles;
  int64_t nd = 0;
  for (int j = 0; j <= c->L; ++j) nd += tmf::total(c->s2[j]);
  q[3] = nd;
  return TMF_OK;
}

// dims2[L + 1], charges2[sum dims2]
extern "C" int tmf_canon_dims(const tmf_canon *c, int *dims2, int *charges2) {
  int64_t o = 0;
  for (int j = 0; j <= c->L; ++j) {
    dims2[j] = tmf::total(c->s2[j]);
    for (const tmf::Sector &s : c->s2[j])
      for (int i = 0; i < s.count; ++i) charges2[o++] = s.q;
  }
  return TMF_OK;
}

// T0_dev + t0_off[j]: tensor of site j, dims0[j] x 2 x dims0[j+1] row-major.  T2_dev: the right-canonical tensors,
// site j at sum_{i<j} dims2[i] * 2 * dims2[i+1]; S_dev: singular values of bond j at sum_{i<j} dims2[i] (bonds
// 0 .. L-1, un-normalised); inv_dev[j]: 1 / norm of the singular values of bond j (lambda = S * inv).
extern "C" int tmf_canon_run(tmf_canon *c, const double *T0_dev, const int64_t *t0_off, void *work_dev,
                             int64_t work_bytes, double *T2_dev, double *S_dev, double *inv_dev, void *stream) {
  using namespace tmf;
  const int L = c->L;
  const int *qp = c->qp;
  CanonBuffers bf;
  canon_layout(c, work_dev, work_bytes, bf);
  if (!bf.ok) { set_error("tmf_canon_run: workspace too small"); return TMF_ERR_VALUE; }
  std::vector<double *> &Tcur = bf.Tcur, &T1 = bf.T1, &Xcur = bf.Xcur;
  std::vector<std::vector<double *>> &Rb = bf.Rb, &USb = bf.USb;
  std::vector<double *> T2(L);
  std::vector<int> d0(L + 1), d1(L + 1), d2(L + 1);
  for (int j = 0; j <= L; ++j) { d0[j] = total(c->s0[j]); d1[j] = total(c->s1[j]); d2[j] = total(c->s2[j]); }
  int64_t o2 = 0;
  for (int j = 0; j < L; ++j) {
    T2[j] = T2_dev + o2;
    o2 += (int64_t)d2[j] * 2 * d2[j + 1];
  }
  double *scratch = bf.scratch, *nrm = bf.nrm, *inv1 = bf.inv1, *invx = bf.invx;
  unsigned char *jobs_dev = bf.jobs;
  void *gdesc = bf.gdesc;
  const int64_t njobs_total = bf.njobs;
  int rc = memset_dev(work_dev, 0, (size_t)bf.zero_bytes, stream);          // tensor buffers, R, US
  if (rc) return rc;
  rc = memset_dev(T2_dev, 0, sizeof(double) * (size_t)c->t2_doubles, stream);
  if (rc) return rc;
  // ---- all block descriptors of both sweeps, one upload ---------------------------------------------------
  std::vector<unsigned char> jh((size_t)128 * (njobs_total + 1), 0);
  std::vector<int64_t> qr_first(L + 1, 0), svd_first(L + 1, 0);
  int64_t ji = 0;
  std::vector<int> qr_nmax(L + 1, 1), svd_camax(L + 1, 1);
  std::vector<int64_t> qr_need(L + 1, 0), svd_need(L + 1, 0);      // largest matrix footprint of a step (doubles)
  for (int j = 0; j + 1 < L; ++j) {
    qr_first[j] = ji;
    int64_t so = 0;
    const double *src = (j == 0) ? T0_dev + t0_off[0] : Tcur[j];
    const int nb = d0[j + 1], nq = d1[j + 1];
    for (size_t u = 0; u < c->s1[j + 1].size(); ++u, ++ji) {
      const Sector &col = c->s0[j + 1][c->src1[j + 1][u]], &nw = c->s1[j + 1][u];
      BlockQrJob &q = *reinterpret_cast<BlockQrJob *>(jh.data() + 128 * ji);
      int m = 0;
      for (int p = 0; p < 2; ++p) {
        const Sector *r = find_sector(c->s1[j], col.q - qp[p]);
        q.m[p] = r ? r->count : 0;
        const int a0 = r ? r->start : 0;
        q.src[p] = src + ((int64_t)a0 * 2 + p) * nb + col.start;
        q.dst[p] = T1[j] + ((int64_t)a0 * 2 + p) * nq + nw.start;
        m += q.m[p];
      }
      q.rs = 2 * (int64_t)nb; q.rd = 2 * (int64_t)nq;
      q.n = col.count; q.k = nw.count;
      q.R = Rb[j + 1][u];
      q.work = scratch + so;
      so += align256(8 * (int64_t)m * col.count) / 8;
      q.nrm2 = nrm + ji;
      q.scale = (j == 0) ? nullptr : inv1 + j;          // 1 / ||R of bond j||
      qr_nmax[j] = std::max(qr_nmax[j], col.count);
      qr_need[j] = std::max<int64_t>(qr_need[j], (int64_t)m * col.count + (int64_t)nw.count * col.count);
    }
  }
  for (int j = L - 1; j >= 0; --j) {
    svd_first[j] = ji;
    int64_t so = 0;
    const double *src = (j == L - 1) ? ((L == 1) ? T0_dev + t0_off[0] : Tcur[j]) : Xcur[j];
    const int nbx = d2[j + 1];
    for (size_t u = 0; u < c->s2[j].size(); ++u, ++ji) {
      const Sector &row = c->s1[j][c->src2[j][u]], &nw = c->s2[j][u];
      BlockSvdJob &q = *reinterpret_cast<BlockSvdJob *>(jh.data() + 128 * ji);
      int n = 0;
      for (int p = 0; p < 2; ++p) {
        const Sector *s = find_sector(c->s2[j + 1], row.q + qp[p]);
        q.n[p] = s ? s->count : 0;
        const int b0 = s ? s->start : 0;
        q.src[p] = src + ((int64_t)row.start * 2 + p) * nbx + b0;
        q.dst[p] = T2[j] + ((int64_t)nw.start * 2 + p) * nbx + b0;
        n += q.n[p];
      }
      q.rs = q.rd = 2 * (int64_t)nbx;
      q.m = row.count; q.k = nw.count;
      q.US = USb[j][u];
      int64_t soff = 0;
      for (int i = 0; i < j; ++i) soff += d2[i];
      q.S = S_dev + soff + nw.start;
      q.work = scratch + so;
      so += align256(8 * svd_work_doubles(row.count, n)) / 8;
      q.nrm2 = nrm + ji;
      q.scale = (j == L - 1) ? invx : inv_dev + j + 1;   // 1 / ||last tensor||, 1 / ||S of bond j+1||
      svd_camax[j] = std::max(svd_camax[j], std::min(row.count, n));
      {
        const int64_t ra = std::max(row.count, n), ca = std::min(row.count, n);
        svd_need[j] = std::max<int64_t>(svd_need[j], ra * ca + 2 * ca * ca);
      }
    }
  }
  rc = copy_h2d(jobs_dev, jh.data(), jh.size(), stream);
  if (rc) return rc;

  std::vector<tmf_gemm_job> g;
  auto zero_job = [](tmf_gemm_job &t) { std::memset(&t, 0, sizeof(t)); t.alpha = 1.0; t.beta = 0.0; };
  // ---- QR sweep ----------------------------------------------------------------------------------------------
  for (int j = 0; j + 1 < L; ++j) {
    const int nj = (int)c->s1[j + 1].size();
    if (nj > 0) {
      const int64_t scr = (int64_t)block_qr_scratch(qr_nmax[j]);
      const int64_t mat = std::max<int64_t>(0, std::min<int64_t>(qr_need[j], (CANON_SMEM_BYTES - scr) / 8));
      rc = launch_t("canon_qr", block_qr_kernel, nj, 256, (size_t)(scr + 8 * mat), stream,
                    reinterpret_cast<const BlockQrJob *>(jobs_dev + 128 * qr_first[j]), qr_nmax[j], (int)mat);
      if (rc) return rc;
    }
    rc = launch_t("canon_norm", inv_norm_kernel, 1, 256, 256 * sizeof(double), stream,
                  (const double *)(nrm + qr_first[j]), (int64_t)nj, 0, inv1 + j + 1);
    if (rc) return rc;
    // T_{j+1}[new sector rows, p, sector(q + qp[p]) of bond j+2] = R . T0_{j+1}[old sector rows, p, ...]
    g.clear();
    const double *Tn = T0_dev + t0_off[j + 1];
    const int nb2 = d0[j + 2];
    for (int u = 0; u < nj; ++u) {
      const Sector &col = c->s0[j + 1][c->src1[j + 1][u]], &nw = c->s1[j + 1][u];
      for (int p = 0; p < 2; ++p) {
        const Sector *cs = find_sector(c->s0[j + 2], col.q + qp[p]);
        if (!cs) continue;
        tmf_gemm_job t;
        zero_job(t);
        t.A = Tn + ((int64_t)col.start * 2 + p) * nb2 + cs->start; t.lda = 2 * nb2; t.transA = 0;
        t.B = Rb[j + 1][u]; t.ldb = nw.count; t.transB = 1;
        t.C = Tcur[j + 1] + ((int64_t)nw.start * 2 + p) * nb2 + cs->start; t.ldc = 2 * nb2;
        t.M = cs->count; t.N = nw.count; t.K = col.count;
        g.push_back(t);
      }
    }
    if (!g.empty()) {
      rc = gemm_grouped(g.data(), (int)g.size(), gdesc, stream);
      if (rc) return rc;
    }
  }
  // ---- SVD sweep -----------------------------------------------------------------------------------------------
  {
    const double *last = (L == 1) ? T0_dev + t0_off[0] : Tcur[L - 1];
    rc = launch_t("canon_norm", inv_norm_kernel, 1, 256, 256 * sizeof(double), stream, last,
                  (int64_t)d1[L - 1] * 2 * d0[L], 1, invx);
    if (rc) return rc;
  }
  for (int j = L - 1; j >= 0; --j) {
    const int nj = (int)c->s2[j].size();
    if (nj > 0) {
      const int64_t scr = ((int64_t)block_svd_scratch(svd_camax[j]) + 15) & ~int64_t(15);
      const int64_t mat = std::max<int64_t>(0, std::min<int64_t>(svd_need[j], (CANON_SMEM_BYTES - scr) / 8));
      rc = launch_t("canon_svd", block_svd_kernel, nj, 256, (size_t)(scr + 8 * mat), stream,
                    reinterpret_cast<const BlockSvdJob *>(jobs_dev + 128 * svd_first[j]), svd_camax[j], (int)mat);
      if (rc) return rc;
    }
    rc = launch_t("canon_norm", inv_norm_kernel, 1, 256, 256 * sizeof(double), stream,
                  (const double *)(nrm + svd_first[j]), (int64_t)nj, 0, inv_dev + j);
    if (rc) return rc;
    if (j == 0) break;
    // X_{j-1}[sector(q - qp[p]) rows of bond j-1, p, new sector cols] = T1_{j-1}[.., p, old sector cols] . (U S)
    g.clear();
    const int nq = d1[j], nd = d2[j];
    for (int u = 0; u < nj; ++u) {
      const Sector &row = c->s1[j][c->src2[j][u]], &nw = c->s2[j][u];
      for (int p = 0; p < 2; ++p) {
        const Sector *rs = find_sector(c->s1[j - 1], row.q - qp[p]);
        if (!rs) continue;
        tmf_gemm_job t;
        zero_job(t);
        t.A = USb[j][u]; t.lda = row.count; t.transA = 1;
        t.B = T1[j - 1] + ((int64_t)rs->start * 2 + p) * nq + row.start; t.ldb = 2 * nq; t.transB = 0;
        t.C = Xcur[j - 1] + ((int64_t)rs->start * 2 + p) * nd + nw.start; t.ldc = 2 * nd;
        t.M = nw.count; t.N = rs->count; t.K = row.count;
        g.push_back(t);
      }
    }
    if (!g.empty()) {
      rc = gemm_grouped(g.data(), (int)g.size(), gdesc, stream);
      if (rc) return rc;
    }
  }
  return TMF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// K15: Procrustes gauge fixing of the iMPS conversion on the device
// ---------------------------------------------------------------------------------------------------------------
extern "C" int64_t tmf_procrustes_workspace(const tmf_procrustes_job *jobs_host, int njobs) {
  using namespace tmf;
  int64_t w = align256(128 * (int64_t)std::max(njobs, 1)) + align256(sizeof(ProcrustesFinish) * (int64_t)std::max(njobs, 1));
  for (int i = 0; i < njobs; ++i) {
    const int m = jobs_host[i].m, n = jobs_host[i].n, k = std::min(m, n);
    w += align256(8 * ((int64_t)k * n)) + align256(8 * ((int64_t)m * k)) + align256(8 * (int64_t)(k + 2)) +
         align256(8 * svd_work_doubles(m, n));
  }
  return w + 4096;
}

extern "C" int tmf_procrustes_blocks(const tmf_procrustes_job *jobs_host, int njobs, void *work_dev,
                                     int64_t work_bytes, void *stream) {
  using namespace tmf;
  if (njobs <= 0) return TMF_OK;
  Arena ar(work_dev, work_bytes);
  unsigned char *sj_dev = ar.take<unsigned char>(128 * (int64_t)njobs);
  unsigned char *fj_dev = ar.take<unsigned char>((int64_t)sizeof(ProcrustesFinish) * njobs);
  std::vector<BlockSvdJob> sj(njobs);
  std::vector<ProcrustesFinish> fj(njobs);
  int camax = 1;
  int64_t need = 0;
  for (int i = 0; i < njobs; ++i) {
    const tmf_procrustes_job &q = jobs_host[i];
    const int m = q.m, n = q.n, k = std::min(m, n);
    if (m <= 0 || n <= 0) { set_error("tmf_procrustes_blocks: empty block"); return TMF_ERR_VALUE; }
    if (k > CANON_MAX) { set_error("tmf_procrustes_blocks: charge block larger than 160"); return TMF_ERR_VALUE; }
    BlockSvdJob &s = sj[i];
    std::memset(&s, 0, sizeof(s));
    double *Vh = ar.take<double>((int64_t)k * n), *U = ar.take<double>((int64_t)m * k), *S = ar.take<double>(k + 2);
    s.src[0] = q.C; s.src[1] = q.C; s.rs = q.ldc;
    s.dst[0] = Vh; s.dst[1] = Vh; s.rd = n;
    s.US = U; s.S = S; s.nrm2 = S + k;
    s.work = ar.take<double>(svd_work_doubles(m, n));
    s.m = m; s.n[0] = n; s.n[1] = 0; s.k = k;
    s.csq = q.sk; s.want_u = 1;
    ProcrustesFinish &f = fj[i];
    f.C = q.C; f.sk = q.sk; f.U = U; f.Vh = Vh; f.S = S; f.R = q.R; f.metrics = q.metrics;
    f.ldc = q.ldc; f.ldr = q.ldr; f.m = m; f.n = n; f.k = k; f.pad_ = 0;
    camax = std::max(camax, k);
    const int64_t ra = std::max(m, n);
    need = std::max<int64_t>(need, ra * k + 2 * (int64_t)k * k);
  }
  if (!ar.ok()) { set_error("tmf_procrustes_blocks: workspace too small"); return TMF_ERR_VALUE; }
  int rc = copy_h2d(sj_dev, sj.data(), sizeof(BlockSvdJob) * (size_t)njobs, stream);
  if (rc) return rc;
  rc = copy_h2d(fj_dev, fj.data(), sizeof(ProcrustesFinish) * (size_t)njobs, stream);
  if (rc) return rc;
  const int64_t scr = ((int64_t)block_svd_scratch(camax) + 15) & ~int64_t(15);
  const int64_t mat = std::max<int64_t>(0, std::min<int64_t>(need, (CANON_SMEM_BYTES - scr) / 8));
  rc = launch_t("procrustes_svd", block_svd_kernel, njobs, 256, (size_t)(scr + 8 * mat), stream,
                reinterpret_cast<const BlockSvdJob *>(sj_dev), camax, (int)mat);
  if (rc) return rc;
  return launch_t("procrustes", procrustes_finish_kernel, njobs, 256, 2 * 256 * sizeof(double), stream,
                  reinterpret_cast<const ProcrustesFinish *>(fj_dev));
}
