// K14: Gutzwiller projection of pairs of fermion sites onto spin-1/2 sites -- one fused launch for a whole MPS.
//
// reference: gutzwiller.py:227 / :409 (mps.group_sites(2): theta = B_{2i} . B_{2i+1}, a dense contraction inside
// TeNPy) followed by gutzwiller.py:242 / :424 (iproject with the occupation / parity masks, which throws most of
// theta away) and the re-labelling of the legs (:244, :437-441).  Here only the charge chains
//     (q_L, p_a) -> q_m -> (p_b, q_R)
// that survive the three masks are ever multiplied.  A job is one such chain,
//     out[vL, s, vR] = row_scale[vL] * sum_m A[vL, m] * k_scale[m] * B[m, vR] * col_scale[vR],
// with A / B addressed *in place* inside the block-sparse fermion tensors where the conversion left them in HBM
// (element strides cover both storage orders: left-canonical tensors keep [vL, vR] rows, right-canonical ones
// [vR, vL]), the Schmidt values of the orthogonality centre folded in as one of the three scalings, and the result
// written directly at its position in the dense spin-site tensor T[vL, s, vR] (row stride so_i = 2 chi_R; the
// launch zero-fills the rest).  Real (Slater) and complex (complex Slater, Pfaffian) tensors share the kernel.
//
// CTA = 64 x 64 output tile of one job, 256 threads, 4 x 4 accumulators per thread, operand tiles of depth 16
// staged in shared memory.  The products are small (blocks of a few tens to a few hundred rows, ~10 chains per
// spin site), the launch is bound by reading the fermion blocks once.
#include "cplx.hpp"

namespace tmf {

static_assert(sizeof(tmf_gutz_job) == 128, "descriptor must be 128 bytes");
constexpr int GT = 64, GK = 16;

TMF_HD void gz_zero(double &a) { a = 0.0; }
TMF_HD void gz_zero(cplx &a) { a = cmake(0.0); }
TMF_HD double gz_fma(double a, double b, double c) { return a * b + c; }
TMF_HD cplx gz_fma(cplx a, cplx b, cplx c) { return cadd(c, cmul(a, b)); }
TMF_HD double gz_scale(double a, double s) { return a * s; }
TMF_HD cplx gz_scale(cplx a, double s) { return cscale(a, s); }

template <class T>
TMF_DEVICE T gutz_a(const tmf_gutz_job &j, int i, int k) {
  T z;
  gz_zero(z);
  if (i >= j.m || k >= j.k) return z;
  T v = static_cast<const T *>(j.A)[i * j.sa_i + k * j.sa_k];
  return j.k_scale ? gz_scale(v, j.k_scale[k]) : v;
}
template <class T>
TMF_DEVICE T gutz_b(const tmf_gutz_job &j, int k, int n) {
  T z;
  gz_zero(z);
  if (k >= j.k || n >= j.n) return z;
  return static_cast<const T *>(j.B)[k * j.sb_k + n * j.sb_n];
}
template <class T>
TMF_DEVICE void gutz_store(const tmf_gutz_job &j, int i, int n, T acc) {
  if (i >= j.m || n >= j.n) return;
  double s = 1.0;
  if (j.row_scale) s *= j.row_scale[i];
  if (j.col_scale) s *= j.col_scale[n];
  static_cast<T *>(j.out)[i * j.so_i + n] = gz_scale(acc, s);
}

// tiles: 4 ints per CTA = (job, first row, first column, 0)
template <class T>
TMF_GLOBAL_LB(256, 2) gutz_pair_kernel(const tmf_gutz_job *jobs, const int *tiles) {
  const int *tl = tiles + 4 * (int64_t)BLOCK_ID;
  const tmf_gutz_job &j = jobs[tl[0]];
  const int m0 = tl[1], n0 = tl[2];
#if defined(TMF_HOSTSIM)
  for (int i = m0; i < m0 + GT && i < j.m; ++i)
    for (int n = n0; n < n0 + GT && n < j.n; ++n) {
      T acc;
      gz_zero(acc);
      for (int k = 0; k < j.k; ++k) acc = gz_fma(gutz_a<T>(j, i, k), gutz_b<T>(j, k, n), acc);
      gutz_store<T>(j, i, n, acc);
    }
#else
  __shared__ T As[GK][GT + 1], Bs[GK][GT + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  T acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) gz_zero(acc[a][b]);
  const bool a_kfast = (j.sa_k == 1), b_nfast = (j.sb_n == 1);
  for (int k0 = 0; k0 < j.k; k0 += GK) {
#pragma unroll
    for (int e = tid; e < GT * GK; e += 256) {
      const int ia = a_kfast ? (e >> 4) : (e & 63), ka = a_kfast ? (e & 15) : (e >> 6);
      As[ka][ia] = gutz_a<T>(j, m0 + ia, k0 + ka);
      const int nb = b_nfast ? (e & 63) : (e >> 4), kb = b_nfast ? (e >> 6) : (e & 15);
      Bs[kb][nb] = gutz_b<T>(j, k0 + kb, n0 + nb);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      T fa[4], fb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) fa[a] = As[kk][ty + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) fb[b] = Bs[kk][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = gz_fma(fa[a], fb[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) gutz_store<T>(j, m0 + ty + 16 * a, n0 + tx + 16 * b, acc[a][b]);
#endif
}

static int64_t gutz_tiles(const tmf_gutz_job *jobs, int njobs) {
  int64_t t = 0;
  for (int i = 0; i < njobs; ++i)
    t += (int64_t)((jobs[i].m + GT - 1) / GT) * ((jobs[i].n + GT - 1) / GT);
  return t;
}
}  // namespace tmf

extern "C" int64_t tmf_gutz_desc_bytes(const tmf_gutz_job *jobs_host, int njobs) {
  return tmf::align256((int64_t)sizeof(tmf_gutz_job) * std::max(njobs, 1)) +
         tmf::align256(16 * std::max<int64_t>(tmf::gutz_tiles(jobs_host, njobs), 1)) + 256;
}

extern "C" int tmf_gutzwiller_project(const tmf_gutz_job *jobs_host, int njobs, int cplx_flag, void *out_dev,
                                      int64_t out_bytes, void *desc_dev, void *stream) {
  using namespace tmf;
  int rc = memset_dev(out_dev, 0, (size_t)out_bytes, stream);
  if (rc || njobs <= 0) return rc;
  for (int i = 0; i < njobs; ++i) {
    const tmf_gutz_job &j = jobs_host[i];
    if (j.m <= 0 || j.n <= 0 || j.k < 0 || !j.A || !j.B || !j.out) {
      set_error("tmf_gutzwiller_project: malformed job");
      return TMF_ERR_VALUE;
    }
  }
  const int64_t ntiles = gutz_tiles(jobs_host, njobs);
  if (ntiles > 0x7fffffff) { set_error("tmf_gutzwiller_project: too many tiles"); return TMF_ERR_VALUE; }
  std::vector<int> tiles((size_t)ntiles * 4);
  int64_t t = 0;
  for (int i = 0; i < njobs; ++i)
    for (int m0 = 0; m0 < jobs_host[i].m; m0 += GT)
      for (int n0 = 0; n0 < jobs_host[i].n; n0 += GT) {
        tiles[4 * t] = i; tiles[4 * t + 1] = m0; tiles[4 * t + 2] = n0; tiles[4 * t + 3] = 0;
        ++t;
      }
  unsigned char *base = static_cast<unsigned char *>(desc_dev);
  tmf_gutz_job *jd = reinterpret_cast<tmf_gutz_job *>(base);
  int *td = reinterpret_cast<int *>(base + align256((int64_t)sizeof(tmf_gutz_job) * njobs));
  rc = copy_h2d(jd, jobs_host, sizeof(tmf_gutz_job) * (size_t)njobs, stream);
  if (rc) return rc;
  rc = copy_h2d(td, tiles.data(), sizeof(int) * tiles.size(), stream);
  if (rc) return rc;
  if (cplx_flag)
    return launch_t("gutzwiller", gutz_pair_kernel<cplx>, (int)ntiles, 256, 0, stream, (const tmf_gutz_job *)jd,
                    (const int *)td);
  return launch_t("gutzwiller", gutz_pair_kernel<double>, (int)ntiles, 256, 0, stream, (const tmf_gutz_job *)jd,
                  (const int *)td);
}
