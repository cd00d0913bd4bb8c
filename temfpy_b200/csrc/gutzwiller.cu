// K14: Gutzwiller projection of a pair of fermion sites onto one spin-1/2 site.
//
// reference: gutzwiller.py:227 / :409 (mps.group_sites(2): theta = B_{2i} . B_{2i+1}) followed by
// gutzwiller.py:242 / :424 (iproject with the occupation / parity masks).  All arithmetic of the
// reference happens inside TeNPy; here only the charge-block chains that survive the three masks
// are multiplied, each as one job of the grouped DMMA GEMM, and the result is written directly in
// the spin-site layout.  Blocks are dense row-major (m x k)(k x n) -> (m x n).
#include "cta.hpp"

namespace tmf {
int gemm_grouped(const tmf_gemm_job *jobs, int njobs, void *desc_dev, void *stream);
}

extern "C" int tmf_gutzwiller_site(const tmf_gutz_job *jobs_host, int njobs, void *desc_dev,
                                   void *stream) {
  if (njobs <= 0) return TMF_OK;
  std::vector<tmf_gemm_job> g(njobs);
  for (int i = 0; i < njobs; ++i) {
    const tmf_gutz_job &q = jobs_host[i];
    tmf_gemm_job &j = g[i];
    std::memset(&j, 0, sizeof(j));
    // row-major out (m x n) = A (m x k) B (k x n)  <=>  column-major out^T (n x m) = B^T A^T
    j.A = q.B; j.lda = q.n; j.transA = 0;   // B^T is n x k column-major with ld n
    j.B = q.A; j.ldb = q.k; j.transB = 0;   // A^T is k x m column-major with ld k
    j.C = q.out; j.ldc = q.n;
    j.M = q.n; j.N = q.m; j.K = q.k;
    j.alpha = 1.0; j.beta = 0.0;
  }
  return tmf::gemm_grouped(g.data(), njobs, desc_dev, stream);
}
