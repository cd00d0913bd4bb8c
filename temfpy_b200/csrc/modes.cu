// K3/K4: batched Schmidt-mode extraction for every (bond, side) job of a chain.
//
// Reference behaviour (slater.py:324-375): eigh of the diagonal block A = C[:x,:x] / C[x:,x:],
// split at cutoff = svd_min^2 into filled / entangled / empty eigenvectors.  Only the k (~20)
// entangled eigenpairs and *any* orthonormal basis of the filled eigenspace enter the MPS tensors
// (never-occupied orbitals are dropped by _select_orbitals, slater.py:792-797; the filled block
// only appears through det / Schur complement, which are basis independent up to a sign that
// cancels between the two tensors sharing the bond).  For a Slater determinant C is a projector,
// hence A(1-A) = B B^T with B the off-diagonal block: the entangled eigenvectors of A are exactly
// the left singular vectors of B with non-zero singular value s = sqrt(e(1-e)).  We therefore
//   1. sketch the range of B with a fixed pseudo-random test matrix (GEMM),
//   2. orthonormalise (block Gram-Schmidt: DMMA GEMMs between panels + in-panel MGS2),
//   3. reduce W = Q^T B to r x r (second QR) and take its SVD by one-sided Jacobi in smem
//      (un-squared, so the weakest entangled modes s ~ svd_min keep ~1e-9 relative accuracy),
//   4. Rayleigh-Ritz A on the selected subspace -> eigenvalues e and eigenvectors,
//   5. pivoted Cholesky of the remaining projector A - U diag(e) U^T -> filled-space basis.
// Blocks with n <= 64 are diagonalised directly by Jacobi.  All bonds of the chain run in the
// same launches (grids of hundreds of CTAs), descriptors are uploaded once up front.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <map>
#include <mutex>
#include "modes_kernels.cuh"

namespace {
struct ModesTimer {   // TMF_DEBUG_TIMING=1 prints host-side stage times to stderr
  bool on;
  std::chrono::steady_clock::time_point t;
  ModesTimer() : on(std::getenv("TMF_DEBUG_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
  void lap(const char *what, int n) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[tmf timing] %-28s %8.3f ms  (%d jobs)\n", what, std::chrono::duration<double, std::milli>(now - t).count(), n);
    t = now;
  }
};
}  // namespace

namespace tmf {

int gemm_grouped(const tmf_gemm_job *jobs, int njobs, void *desc_dev, void *stream);
int64_t gemm_desc_bytes(int njobs);

// deterministic test matrix: Omega[row + col * ld] in (-1, 1) from a hash of (row, col)
TMF_GLOBAL omega_kernel(double *om, int rows, int cols) {
  const int64_t total = (int64_t)rows * cols;
  PAR_FOR(t, 1024) {  // each CTA fills 1024 consecutive entries
    int64_t g = (int64_t)BLOCK_ID * 1024 + t;
    if (g < total) {
      uint64_t z = (uint64_t)g * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z = z ^ (z >> 31);
      om[g] = ((double)(z >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
    }
  }
}

// Range sketch Y_x = B_x Omega for every bond x at once.  With the same test matrix for all bonds
// (rows of Omega indexed by the global site) the sketches of consecutive bonds are prefix / suffix sums,
//     right block at x:  Y_x[g - x, c] = sum_{j <  x} C[g, j] Omega[j, c]   (g >= x)
//     left  block at x:  Y_x[g,     c] = sum_{j >= x} C[g, j] Omega[j, c]   (g <  x)
// so one pass over C per sketch column (L^2 r multiply-adds in total) replaces the batched GEMM
// (sum_x 2 n m r, ~2e10 flop at L = 1024).  Thread = (site g, sketch column c), lanes along g: the reads of
// C (symmetric, so C[g, j] = C[j, g]) and the writes of Y coalesce.
struct SketchEntry {
  double *Y;   // n x rr sketch of the job at this bond (nullptr: no job)
  int n, rr;
};
TMF_GLOBAL sketch_scan_kernel(const double *C, int ldc, int L, const double *Om, int r, const SketchEntry *tabR,
                              const SketchEntry *tabL, int xloR, int xhiR, int xloL, int xhiL) {
  const int gblocks = (L + 31) / 32;
  const int gb = BLOCK_ID % gblocks, cb = BLOCK_ID / gblocks;
  PAR_FOR(t, 256) {
    const int g = gb * 32 + (t & 31), c = cb * 8 + (t >> 5);
    if (g < L && c < r) {
      const double *om = Om + (int64_t)c * L;
      if (xhiR >= xloR) {
        double s = 0.0;
        const int jend = g < xhiR ? g : xhiR;
        for (int j = 0; j <= jend; ++j) {
          if (j >= xloR) {
            const SketchEntry e = tabR[j];
            if (e.Y != nullptr && c < e.rr) e.Y[(int64_t)c * e.n + (g - j)] = s;
          }
          s += C[(int64_t)j * ldc + g] * om[j];
        }
      }
      if (xhiL >= xloL) {
        double s = 0.0;
        const int jbeg = (g + 1 > xloL) ? g + 1 : xloL;
        for (int j = L - 1; j >= jbeg; --j) {
          s += C[(int64_t)j * ldc + g] * om[j];
          if (j <= xhiL) {
            const SketchEntry e = tabL[j];
            if (e.Y != nullptr && c < e.rr) e.Y[(int64_t)c * e.n + g] = s;
          }
        }
      }
    }
  }
}

#if defined(TMF_HOSTSIM)
static int copy_d2d(void *dst, const void *src, size_t bytes, void *) {
  std::memcpy(dst, src, bytes);
  return TMF_OK;
}
#else
static int copy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
  if (!bytes) return TMF_OK;
  return check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream),
                    "cudaMemcpyAsync D2D");
}
#endif

namespace {

struct BigJob {
  int job;        // index into the caller's job list
  int n, m, rr, side;
  const double *A, *B, *Om;
  double *Y, *Wt, *Wt0, *coef, *Rw, *Jsel, *Jwork, *U0, *AU, *TE, *Zsel, *norm0y, *norm0w, *eside;
  int *k0, *nzy, *nzw;
  double *V;
  double *e;
  int *info;
};

// Host-side blob of descriptors mirrored at a device address (single upload).
struct Blob {
  std::vector<unsigned char> host;
  unsigned char *dev;
  explicit Blob(void *d) : dev(static_cast<unsigned char *>(d)) {}
  template <class T>
  const T *add(const std::vector<T> &v) {
    size_t off = (host.size() + 255) & ~size_t(255);
    host.resize(off + sizeof(T) * v.size());
    if (!v.empty()) std::memcpy(host.data() + off, v.data(), sizeof(T) * v.size());
    return reinterpret_cast<const T *>(dev + off);
  }
};

struct GemmLaunch {
  const tmf_gemm_job *jobs;
  const int *prefix;
  int njobs, ntiles;
};
GemmLaunch add_gemm(Blob &blob, const std::vector<tmf_gemm_job> &jobs) {
  std::vector<int> prefix(jobs.size() + 1, 0);
  for (size_t i = 0; i < jobs.size(); ++i) {
    int tm = (jobs[i].M + 63) / 64, tn = (jobs[i].N + 63) / 64;
    if (jobs[i].M <= 0 || jobs[i].N <= 0 || jobs[i].K < 0) tm = tn = 0;
    prefix[i + 1] = prefix[i] + tm * tn;
  }
  GemmLaunch g;
  g.jobs = blob.add(jobs);
  g.prefix = blob.add(prefix);
  g.njobs = (int)jobs.size();
  g.ntiles = prefix.back();
  return g;
}

tmf_gemm_job mk_gemm(const double *A, int lda, int transA, const double *B, int ldb, int transB,
                     double *C, int ldc, int M, int N, int K, double alpha = 1.0, double beta = 0.0) {
  tmf_gemm_job j;
  std::memset(&j, 0, sizeof(j));
  j.A = A; j.B = B; j.C = C;
  j.M = M; j.N = N; j.K = K;
  j.lda = lda; j.ldb = ldb; j.ldc = ldc;
  j.transA = transA; j.transB = transB;
  j.alpha = alpha; j.beta = beta;
  return j;
}

}  // namespace

// defined in gemm.cu (kernel symbol shared through this launcher)
int gemm_launch_uploaded(const tmf_gemm_job *jobs_dev, const int *prefix_dev, int njobs, int ntiles,
                         void *stream, const char *tag);
int gemm_launch_uploaded_tma(const tmf_gemm_job *jobs_dev, const int *prefix_dev, int njobs, int ntiles,
                             void *stream, const char *tag, const double *Cmat, int L, int ldc);

static int64_t big_job_doubles(int n, int m, int rr) {
  // Y, Wt, Wt0, coef, Rw, Jsel, U0, AU, TE, Zsel, norm0y, norm0w, eside (+ ints)
  int64_t d = (int64_t)n * rr + 2 * (int64_t)m * rr + (int64_t)rr * PANEL_W + 5 * (int64_t)rr * rr +
              2 * (int64_t)n * rr + 2 * rr + TMF_MAX_MODES + 16;
  return d + 15 * 32;  // slack for 256-byte alignment of every sub-buffer
}

static void job_geometry(int L, int x, int side, int &n, int &m) {
  if (side == TMF_SIDE_L) { n = x; m = L - x; }
  else { n = L - x; m = x; }
}

}  // namespace tmf

namespace tmf {
// fork / join of an auxiliary stream around independent work (events only; nothing synchronises the host)
struct ForkedStream {
#if defined(TMF_HOSTSIM)
  void *open(void *parent) { return parent; }
  int join(void *) { return TMF_OK; }
#else
  // one auxiliary stream + event per parent stream, created once and kept (stream creation / destruction
  // takes context-wide locks: per-call streams stalled the launches of the other pipeline threads)
  struct Aux { cudaStream_t s = nullptr; cudaEvent_t ev = nullptr; };
  Aux aux;
  static Aux lookup(void *parent) {
    static std::mutex mu;
    static std::map<void *, Aux> cache;
    std::lock_guard<std::mutex> lk(mu);
    Aux &a = cache[parent];
    if (!a.s) {
      if (cudaStreamCreateWithFlags(&a.s, cudaStreamNonBlocking) != cudaSuccess) { a.s = nullptr; return a; }
      if (cudaEventCreateWithFlags(&a.ev, cudaEventDisableTiming) != cudaSuccess) { a.ev = nullptr; a.s = nullptr; }
    }
    return a;
  }
  void *open(void *parent) {
    aux = lookup(parent);
    if (!aux.s) return parent;
    cudaEventRecord(aux.ev, (cudaStream_t)parent);
    cudaStreamWaitEvent(aux.s, aux.ev, 0);
    return aux.s;
  }
  int join(void *parent) {
    if (!aux.s) return TMF_OK;
    cudaEventRecord(aux.ev, aux.s);
    cudaError_t e = cudaStreamWaitEvent((cudaStream_t)parent, aux.ev, 0);
    return check_cuda(e, "stream join");
  }
#endif
};
}  // namespace tmf

extern "C" int64_t tmf_slater_modes_workspace(int L, int njobs, const int *job_x,
                                              const int *job_side, int r_sketch) {
  using namespace tmf;
  int64_t bytes = align256((int64_t)L * r_sketch * 8);  // Omega
  int nbig = 0, nsmall = 0;
  for (int j = 0; j < njobs; ++j) {
    int n, m;
    job_geometry(L, job_x[j], job_side[j], n, m);
    if (n > std::min(small_n(true), small_n(false))) {   // (upper bound of both forms)
      int rr = std::min(r_sketch, std::min(n, m));
      bytes += align256(big_job_doubles(n, m, rr) * 8);
      ++nbig;
    } else {
      ++nsmall;
    }
  }
  // descriptor blob: ~ (14 + 8 * panels) launches of nbig descriptors of 128 bytes
  const int panels = (r_sketch + PANEL_W - 1) / PANEL_W;
  bytes += align256((int64_t)(nbig + 1) * 128 * (16 + 10 * panels) + (int64_t)(nsmall + 1) * 64 + (int64_t)(L + 1) * 32 + 65536);
  bytes += align256((int64_t)(njobs + 2) * 64);   // edge-vector descriptors (nested mode)
  bytes += align256((int64_t)(njobs + 2) * 32);   // complex pairing descriptors (embedded complex matrices)
  return bytes;
}

static int modes_impl(const double *C_dev, int L, int ldc, int njobs,
                      const int *job_x, const int *job_side, double cutoff,
                      int r_sketch, const int64_t *v_off, double *V_dev,
                      double *e_dev, int *info_dev, void *work_dev,
                      int64_t work_bytes, void *stream, bool nested, double *edge_dev, bool emb = false) {
  using namespace tmf;
  if (r_sketch <= 0 || r_sketch > R_SKETCH_MAX) {
    set_error("r_sketch must be in 1..160");
    return TMF_ERR_VALUE;
  }
  if (work_bytes < tmf_slater_modes_workspace(L, njobs, job_x, job_side, r_sketch)) {
    set_error("tmf_slater_modes_batched: workspace too small");
    return TMF_ERR_VALUE;
  }
  ModesTimer tm;
  int rc = memset_dev(info_dev, 0, sizeof(int) * 4 * (size_t)njobs, stream);
  if (rc) return rc;
  tm.lap("modes: memset", njobs);
  Arena arena(work_dev, work_bytes);
  double *Om = arena.take<double>((int64_t)L * r_sketch);

  std::vector<BigJob> big;
  std::vector<SmallJob> small;
  size_t small_smem = 0;
  for (int j = 0; j < njobs; ++j) {
    int n, m;
    const int x = job_x[j], side = job_side[j];
    if (x < 0 || x > L) { set_error("bond position out of range"); return TMF_ERR_VALUE; }
    job_geometry(L, x, side, n, m);
    const double *A = (side == TMF_SIDE_L) ? C_dev : C_dev + (int64_t)x * ldc + x;
    if (n <= small_n(nested)) {
      SmallJob s;
      s.A = A; s.V = V_dev + v_off[j]; s.e_out = e_dev + (int64_t)j * TMF_MAX_MODES;
      s.info = info_dev + 4 * j; s.n = n; s.lda = ldc; s.side = side; s.pad_ = 0;
      small.push_back(s);
      small_smem = std::max(small_smem, small_smem_bytes(std::max(n, 1)));
      continue;
    }
    BigJob b;
    b.job = j; b.n = n; b.m = m; b.side = side;
    b.rr = std::min(r_sketch, std::min(n, m));
    b.A = A;
    b.B = (side == TMF_SIDE_L) ? C_dev + (int64_t)x * ldc : C_dev + x;
    b.Om = (side == TMF_SIDE_L) ? Om + x : Om;
    b.V = V_dev + v_off[j];
    b.e = e_dev + (int64_t)j * TMF_MAX_MODES;
    b.info = info_dev + 4 * j;
    big.push_back(b);
  }
  // largest blocks first: the one-CTA-per-job kernels take longest on them, so they should not start last
  std::stable_sort(big.begin(), big.end(), [](const BigJob &a, const BigJob &b) { return a.n > b.n; });
  const int nb = (int)big.size();
  // sub-buffers are grouped by kind so that Wt -> Wt0 is one device copy and the counters one memset
  for (auto &b : big) b.Y = arena.take<double>((int64_t)b.n * b.rr);
  unsigned char *wt_begin = reinterpret_cast<unsigned char *>(arena.take<double>(0));
  for (auto &b : big) b.Wt = arena.take<double>((int64_t)b.m * b.rr);
  unsigned char *wt_end = reinterpret_cast<unsigned char *>(arena.take<double>(0));
  const size_t wt_bytes = (size_t)(wt_end - wt_begin);
  for (auto &b : big)
    b.Wt0 = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(b.Wt) + wt_bytes);
  arena.take<unsigned char>((int64_t)wt_bytes);
  for (auto &b : big) {
    const int rr = b.rr;
    b.coef = arena.take<double>((int64_t)rr * PANEL_W);
    b.Rw = arena.take<double>((int64_t)rr * rr);
    b.Jsel = arena.take<double>((int64_t)rr * rr);
    b.Jwork = arena.take<double>((int64_t)rr * rr);
    b.TE = arena.take<double>((int64_t)rr * rr);
    b.Zsel = arena.take<double>((int64_t)rr * rr);
    b.norm0y = arena.take<double>(rr);
    b.norm0w = arena.take<double>(rr);
    b.eside = arena.take<double>(TMF_MAX_MODES);
  }
  for (auto &b : big) b.U0 = arena.take<double>((int64_t)b.n * b.rr);
  for (auto &b : big) b.AU = arena.take<double>((int64_t)b.n * b.rr);
  int *counters = arena.take<int>(4 * (int64_t)std::max(nb, 1));
  for (int i = 0; i < nb; ++i) {
    big[i].k0 = counters + 4 * i;
    big[i].nzy = big[i].k0 + 1;
    big[i].nzw = big[i].k0 + 2;
  }
  unsigned char *blob_dev = arena.take<unsigned char>(0);
  if (!arena.ok()) { set_error("modes: workspace overflow"); return TMF_ERR_VALUE; }
  Blob blob(blob_dev);

  // ---- build every descriptor list ----------------------------------------------------------
  std::vector<tmf_gemm_job> g;
  std::vector<GemmLaunch> L_sketch(1), L_wt(1), L_rw(1), L_u0(1), L_au(1), L_te(1), L_out(1);
  struct OrthPlan {
    std::vector<GemmLaunch> coef, upd;   // per panel (index 0 unused)
    std::vector<const PanelJob *> panel;    // first round: zero test against the original sketch column norms
    std::vector<const PanelJob *> panel2;   // second round: zero test against the (unit) norms before the re-projection
    std::vector<size_t> panel_smem;
    std::vector<int> panel_n;
    const NormJob *norm;
    size_t norm_smem;
  } orthY, orthW;

  // squared norm (relative to the original sketch column) below which an orthogonalised column is rounding
  // noise: a real direction at the sketch's noise floor (s^2 = 1e-26 s_max^2) keeps ~1e-26, noise is ~1e-31
  constexpr double SKETCH_NOISE2 = 1e-28;
  // "twice is enough": in the second round every column enters with unit norm; one that loses more than half
  // of it in the re-projection was rounding noise inside the span of the earlier panels (its first-round
  // normalisation amplified that noise) and is zeroed -- otherwise Q picks up copies of earlier directions
  // (seen as an O(1) loss of orthogonality for spectra that fall below the noise within one panel)
  constexpr double REPROJECT_KEEP2 = 0.25;
  const double *ones_dev = blob.add(std::vector<double>(PANEL_W, 1.0));
  static const bool proj_gemm = std::getenv("TMF_PROJ_GEMM") != nullptr;   // debugging switch: projections as separate GEMMs
  static const bool use_mgs = std::getenv("TMF_PANEL_MGS") != nullptr;   // debugging switch: column-by-column MGS2 panels + GEMM projections
  auto build_orth = [&](OrthPlan &op, bool forW) {
    int rmax = 0;
    for (auto &b : big) rmax = std::max(rmax, b.rr);
    const int panels = (rmax + PANEL_W - 1) / PANEL_W;
    std::vector<NormJob> nj;
    for (auto &b : big) {
      NormJob q;
      q.Y = forW ? b.Wt : b.Y; q.out = forW ? b.norm0w : b.norm0y;
      q.rows = forW ? b.m : b.n; q.ld = q.rows; q.ncols = b.rr; q.pad_ = 0;
      nj.push_back(q);
    }
    op.norm = blob.add(nj);
    op.norm_smem = sizeof(double) * (size_t)std::max(rmax, 1) * 33;
    op.coef.resize(panels); op.upd.resize(panels); op.panel.resize(panels); op.panel2.resize(panels);
    op.panel_smem.assign(panels, 0); op.panel_n.assign(panels, 0);
    for (int p = 0; p < panels; ++p) {
      std::vector<tmf_gemm_job> gc, gu;
      std::vector<PanelJob> pj;
      size_t smem = 0;
      for (auto &b : big) {
        const int c0 = p * PANEL_W;
        const int w = std::min(PANEL_W, b.rr - c0);
        if (w <= 0) continue;
        double *M = forW ? b.Wt : b.Y;
        const int rows = forW ? b.m : b.n;
        double *P = M + (int64_t)c0 * rows;
        if (c0 > 0 && (use_mgs || proj_gemm)) {
          // coef (c0 x w) = Qprev^T P ;  P -= Qprev coef   (the Cholesky-QR panel kernel does this itself)
          gc.push_back(mk_gemm(M, rows, 1, P, rows, 0, b.coef, c0, c0, w, rows));
          gu.push_back(mk_gemm(M, rows, 0, b.coef, c0, 0, P, rows, rows, w, c0, -1.0, 1.0));
        }
        PanelJob q;
        std::memset(&q, 0, sizeof(q));
        q.P = P; q.norm0 = (forW ? b.norm0w : b.norm0y) + c0; q.nzero = forW ? b.nzw : b.nzy;
        q.Qprev = (c0 > 0 && !use_mgs && !proj_gemm) ? M : nullptr; q.c0 = c0;
        q.rows = rows; q.ld = rows; q.ncols = w;
        q.use_smem = (panel_smem_bytes(rows, w, true) <= 200 * 1024) ? 1 : 0;
        smem = std::max(smem, panel_smem_bytes(rows, w, q.use_smem != 0));
        pj.push_back(q);
      }
      op.coef[p] = add_gemm(blob, gc);
      op.upd[p] = add_gemm(blob, gu);
      op.panel[p] = blob.add(pj);
      for (auto &q2 : pj) q2.norm0 = ones_dev;
      op.panel2[p] = blob.add(pj);
      op.panel_n[p] = (int)pj.size();
      op.panel_smem[p] = smem;
    }
  };

  // 1. Y = B * Omega: prefix / suffix scan over all bonds (the batched GEMM is kept as a debugging switch)
  static const bool sketch_gemm = std::getenv("TMF_SKETCH_GEMM") != nullptr;
  g.clear();
  if (sketch_gemm)
    for (auto &b : big) g.push_back(mk_gemm(b.B, ldc, 0, b.Om, L, 0, b.Y, b.n, b.n, b.rr, b.m));
  L_sketch[0] = add_gemm(blob, g);
  std::vector<SketchEntry> tabR((size_t)L + 1, SketchEntry{nullptr, 0, 0}), tabL((size_t)L + 1, SketchEntry{nullptr, 0, 0});
  int xloR = L + 1, xhiR = -1, xloL = L + 1, xhiL = -1;
  for (auto &b : big) {
    const int x = job_x[b.job];
    if (b.side == TMF_SIDE_R) {
      tabR[x] = SketchEntry{b.Y, b.n, b.rr};
      xloR = std::min(xloR, x); xhiR = std::max(xhiR, x);
    } else {
      tabL[x] = SketchEntry{b.Y, b.n, b.rr};
      xloL = std::min(xloL, x); xhiL = std::max(xhiL, x);
    }
  }
  const SketchEntry *tabR_dev = blob.add(tabR), *tabL_dev = blob.add(tabL);
  build_orth(orthY, false);
  // 4. Wt = B^T Q
  g.clear();
  // (A operands that are k-contiguous sub-blocks of C are flagged for the TMA-staged kernel: pad_[1] / pad_[2] =
  //  row / column of the block's origin inside C)
  auto flag_tma = [&](tmf_gemm_job &j) {
    const int64_t off = j.A - C_dev;
    j.pad_[0] = 1; j.pad_[1] = (int)(off / ldc); j.pad_[2] = (int)(off % ldc);
  };
  for (auto &b : big) {
    g.push_back(mk_gemm(b.B, ldc, 1, b.Y, b.n, 0, b.Wt, b.m, b.m, b.rr, b.n));
    flag_tma(g.back());
  }
  L_wt[0] = add_gemm(blob, g);
  build_orth(orthW, true);
  // 7. Rw = Qw^T Wt0
  g.clear();
  for (auto &b : big) g.push_back(mk_gemm(b.Wt, b.m, 1, b.Wt0, b.m, 0, b.Rw, b.rr, b.rr, b.rr, b.m));
  L_rw[0] = add_gemm(blob, g);
  // 8. svd select
  std::vector<SvdSelJob> sj;
  size_t svd_smem = 0;
  for (auto &b : big) {
    SvdSelJob q;
    q.Rw = b.Rw; q.Jsel = b.Jsel; q.Jwork = b.Jwork; q.k0_out = b.k0; q.info = b.info; q.nzero = b.nzy; q.rr = b.rr; q.complete = (b.rr >= std::min(b.n, b.m)) ? 1 : 0;
    sj.push_back(q);
    svd_smem = std::max(svd_smem, svdsel_smem_bytes(std::max(b.rr, 1)));
  }
  const SvdSelJob *sj_dev = blob.add(sj);
  // 9-11
  g.clear();
  for (auto &b : big) g.push_back(mk_gemm(b.Y, b.n, 0, b.Jsel, b.rr, 0, b.U0, b.n, b.n, b.rr, b.rr));
  L_u0[0] = add_gemm(blob, g);
  g.clear();
  // A U0 with A read as A^T (the diagonal block of C is symmetric): k-contiguous, TMA-staged
  for (auto &b : big) {
    g.push_back(mk_gemm(b.A, ldc, 1, b.U0, b.n, 0, b.AU, b.n, b.n, b.rr, b.n));
    flag_tma(g.back());
  }
  L_au[0] = add_gemm(blob, g);
  g.clear();
  for (auto &b : big) g.push_back(mk_gemm(b.U0, b.n, 1, b.AU, b.n, 0, b.TE, b.rr, b.rr, b.rr, b.n));
  L_te[0] = add_gemm(blob, g);
  // 12. ritz
  std::vector<RitzJob> rj;
  size_t ritz_smem = 0;
  for (auto &b : big) {
    ritz_smem = std::max(ritz_smem, ritz_smem_bytes(std::max(b.rr, 1)));
    RitzJob q;
    q.TE = b.TE; q.Zsel = b.Zsel; q.Jwork = b.Jwork; q.k0 = b.k0; q.e_out = b.e; q.eside_out = b.eside; q.info = b.info;
    q.rr = b.rr; q.side = b.side;
    rj.push_back(q);
  }
  const RitzJob *rj_dev = blob.add(rj);
  // 13. V[:, :rr] = U0 Zsel
  g.clear();
  for (auto &b : big) g.push_back(mk_gemm(b.U0, b.n, 0, b.Zsel, b.rr, 0, b.V, b.n, b.n, b.rr, b.rr));
  L_out[0] = add_gemm(blob, g);
  // 14. pivoted Cholesky
  std::vector<CholJob> cj;
  size_t chol_smem = 0;
  for (auto &b : big) {
    CholJob q;
    q.A = b.A; q.V = b.V; q.w = b.eside; q.info = b.info; q.n = b.n; q.lda = ldc; q.max_cols = b.n; q.pad_ = 0;
    cj.push_back(q);
    chol_smem = std::max(chol_smem, pivchol_smem_bytes(b.n));
  }
  const CholJob *cj_dev = blob.add(cj);
  const SmallJob *small_dev = blob.add(small);
  // nested mode: edge vectors instead of the filled-space bases (no pivoted Cholesky)
  std::vector<EdgeJob> edge_small, edge_big;
  if (nested) {
    auto mk_edge = [&](int j, const double *A, int n, int side) {
      EdgeJob q;
      q.A = A; q.V = V_dev + v_off[j]; q.e_left = e_dev + (int64_t)j * TMF_MAX_MODES; q.info = info_dev + 4 * j;
      q.edge_out = edge_dev + 2 * (int64_t)j; q.n = n; q.lda = ldc; q.side = side; q.emb = emb ? 1 : 0;
      return q;
    };
    for (auto &sj2 : small) {
      const int j = (int)((sj2.info - info_dev) / 4);
      edge_small.push_back(mk_edge(j, sj2.A, sj2.n, sj2.side));
    }
    for (auto &b : big) edge_big.push_back(mk_edge(b.job, b.A, b.n, b.side));
  }
  const EdgeJob *edge_small_dev = blob.add(edge_small), *edge_big_dev = blob.add(edge_big);
  // embedded complex matrix: complex modes out of the real eigenvector pairs, after everything else
  std::vector<PairCJob> pairc;
  if (emb)
    for (int j = 0; j < njobs; ++j) {
      int n, m;
      job_geometry(L, job_x[j], job_side[j], n, m);
      PairCJob q;
      q.V = V_dev + v_off[j]; q.e_left = e_dev + (int64_t)j * TMF_MAX_MODES; q.info = info_dev + 4 * j; q.n_emb = n; q.side = job_side[j];
      pairc.push_back(q);
    }
  const PairCJob *pairc_dev = blob.add(pairc);

  if ((int64_t)(blob_dev - static_cast<unsigned char *>(work_dev)) + (int64_t)blob.host.size() > work_bytes) {
    set_error("modes: descriptor blob does not fit the workspace");
    return TMF_ERR_VALUE;
  }
  tm.lap("modes: build descriptors", njobs);
  rc = copy_h2d(blob_dev, blob.host.data(), blob.host.size(), stream);
  if (rc) return rc;
  tm.lap("modes: upload descriptors", njobs);

  // ---- launches ----------------------------------------------------------------------------
  // The direct solver of the small blocks (<= 65 latency-bound CTAs) runs on a forked stream next to
  // the sketch path of the large blocks and is joined at the end.
  // (8 lanes per Jacobi column pair for n <= 64: 32 pairs = 256 threads)
  static const int small_threads = std::getenv("TMF_SMALL_THREADS") ? std::atoi(std::getenv("TMF_SMALL_THREADS")) : 256;
  ForkedStream fork;
  if (!small.empty()) {
    void *sstream = (nb > 0) ? fork.open(stream) : stream;
    rc = launch_t("small_modes", small_modes_kernel, (int)small.size(), small_threads, small_smem, sstream, small_dev, cutoff);
    if (rc) return rc;
    if (nested) {
      rc = launch_t("edge_vector", edge_vector_kernel, (int)edge_small.size(), 256, edge_smem_bytes(), sstream, edge_small_dev);
      if (rc) return rc;
    }
  }
  if (nb == 0) {
    if (emb) return launch_t("pair_complex", pair_complex_kernel, njobs, 256, pairc_smem_bytes(), stream, pairc_dev);
    return TMF_OK;
  }
  rc = launch_t("omega", omega_kernel, (int)(((int64_t)L * r_sketch + 1023) / 1024), 256, 0, stream, Om, L, r_sketch);
  if (rc) return rc;
  auto run = [&](const GemmLaunch &gl) {
    if (gl.ntiles == 0) return (int)TMF_OK;
    return gemm_launch_uploaded(gl.jobs, gl.prefix, gl.njobs, gl.ntiles, stream, "gemm_modes");
  };
  auto run_tma = [&](const GemmLaunch &gl) {
    if (gl.ntiles == 0) return (int)TMF_OK;
    return gemm_launch_uploaded_tma(gl.jobs, gl.prefix, gl.njobs, gl.ntiles, stream, "gemm_modes", C_dev, L, ldc);
  };
  auto run_orth = [&](OrthPlan &op) -> int {
    int r2 = launch_t("colnorm", colnorm_kernel, nb, 256, op.norm_smem, stream, op.norm);
    if (r2) return r2;
    for (size_t p = 0; p < op.panel.size(); ++p) {
      if (op.panel_n[p] == 0) continue;
      // BCGS2: (project against the previous panels, orthonormalise the panel) twice.  The second
      // round removes what the in-panel normalisation of small columns amplified.
      for (int round = 0; round < 2; ++round) {
        if (p > 0) {
          if ((r2 = run(op.coef[p]))) return r2;
          if ((r2 = run(op.upd[p]))) return r2;
        }
        if (use_mgs)
          r2 = launch_t("panel_mgs2", panel_mgs2_kernel, op.panel_n[p], 256, op.panel_smem[p], stream,
                        round == 0 ? op.panel[p] : op.panel2[p], round == 0 ? SKETCH_NOISE2 : REPROJECT_KEEP2);
        else
          r2 = launch_t("panel_cholqr", panel_cholqr_kernel, op.panel_n[p], 256, panel_cholqr_smem_bytes((int)p * PANEL_W), stream,
                        round == 0 ? op.panel[p] : op.panel2[p], round == 0 ? SKETCH_NOISE2 : REPROJECT_KEEP2);
        if (r2) return r2;
      }
    }
    return TMF_OK;
  };
  rc = memset_dev(counters, 0, sizeof(int) * 4 * (size_t)nb, stream);
  if (rc) return rc;
  if (sketch_gemm) {
    if ((rc = run(L_sketch[0]))) return rc;
  } else {
    rc = launch_t("sketch_scan", sketch_scan_kernel, ((L + 31) / 32) * ((r_sketch + 7) / 8), 256, 0, stream, C_dev, ldc, L,
                  (const double *)Om, r_sketch, tabR_dev, tabL_dev, xloR, xhiR, xloL, xhiL);
    if (rc) return rc;
  }
  if ((rc = run_orth(orthY))) return rc;
  if ((rc = run_tma(L_wt[0]))) return rc;
  rc = copy_d2d(wt_begin + wt_bytes, wt_begin, wt_bytes, stream);
  if (rc) return rc;
  if ((rc = run_orth(orthW))) return rc;
  if ((rc = run(L_rw[0]))) return rc;
  const double thr = cutoff * (1.0 - cutoff);
  static const double sketch_floor = std::getenv("TMF_SKETCH_FLOOR") ? std::atof(std::getenv("TMF_SKETCH_FLOOR")) : 1e-26;
  // one warp per column pair of a Jacobi round (latency-bound: more warps per CTA, not more CTAs)
  static const int jac_env = std::getenv("TMF_JAC_THREADS") ? std::atoi(std::getenv("TMF_JAC_THREADS")) : 0;
  // one group of GW lanes per column pair of a Jacobi round (jacobi_onesided: GW = 8 / 16 / 32 by matrix size)
  const int jac_gw = (r_sketch <= 64) ? 8 : (r_sketch <= 128 ? 16 : 32);
  const int jac_threads = jac_env > 0 ? jac_env : std::min(1024, std::max(128, (jac_gw * ((r_sketch + 1) / 2) + 31) & ~31));
  rc = launch_t("svd_select", svd_select_kernel, nb, jac_threads, svd_smem, stream, sj_dev, thr, sketch_floor);
  if (rc) return rc;
  if ((rc = run(L_u0[0]))) return rc;
  if ((rc = run_tma(L_au[0]))) return rc;
  if ((rc = run(L_te[0]))) return rc;
  rc = launch_t("ritz", ritz_kernel, nb, jac_threads, ritz_smem, stream, rj_dev, cutoff);
  if (rc) return rc;
  if ((rc = run(L_out[0]))) return rc;
  static const int pivchol_threads = std::getenv("TMF_PIVCHOL_THREADS") ? std::atoi(std::getenv("TMF_PIVCHOL_THREADS")) : 1024;
  // rank tolerance of the pivoted Cholesky: the remaining projector still carries the near-empty modes
  // (0 < e < cutoff, not selected as entangled) as eigenvalues up to `cutoff`, so the residual diagonal that
  // ends the factorisation has to sit above them -- and below 1 / n, the smallest pivot of a true rank
  int n_big = 1;
  for (auto &b : big) n_big = std::max(n_big, b.n);
  const double chol_tol = std::min(std::max(1e-8, 30.0 * cutoff), 0.25 / n_big);
  if (nested)
    rc = launch_t("edge_vector", edge_vector_kernel, nb, 256, edge_smem_bytes(), stream, edge_big_dev);
  else
    rc = launch_t("pivchol", pivchol_kernel, nb, pivchol_threads, chol_smem, stream, cj_dev, chol_tol);
  if (rc) return rc;
  tm.lap("modes: launches", njobs);
  rc = fork.join(stream);
  if (rc) return rc;
  if (emb) rc = launch_t("pair_complex", pair_complex_kernel, njobs, 256, pairc_smem_bytes(), stream, pairc_dev);
  return rc;
}

extern "C" int tmf_slater_modes_batched(const double *C_dev, int L, int ldc, int njobs,
                                        const int *job_x, const int *job_side, double cutoff,
                                        int r_sketch, const int64_t *v_off, double *V_dev,
                                        double *e_dev, int *info_dev, void *work_dev,
                                        int64_t work_bytes, void *stream) {
  return modes_impl(C_dev, L, ldc, njobs, job_x, job_side, cutoff, r_sketch, v_off, V_dev, e_dev, info_dev, work_dev,
                    work_bytes, stream, false, nullptr);
}

// Columns a job's V slot must hold: the legacy form stores [entangled | filled basis] (n columns), the
// nested form [entangled | edge vector] (at most min(n, m, r_sketch) Ritz columns + 1; small blocks are
// diagonalised completely: n + 1).
extern "C" int64_t tmf_slater_modes_slot_cols(int L, int x, int side, int r_sketch, int nested) {
  int n, m;
  tmf::job_geometry(L, x, side, n, m);
  if (!nested) return n;
  if (n <= tmf::small_n(true)) return n + 1;
  return std::min(r_sketch, std::min(n, m)) + 1;
}

extern "C" int tmf_slater_modes_nested(const double *C_dev, int L, int ldc, int njobs,
                                       const int *job_x, const int *job_side, double cutoff,
                                       int r_sketch, const int64_t *v_off, double *V_dev,
                                       double *e_dev, int *info_dev, double *edge_dev, void *work_dev,
                                       int64_t work_bytes, void *stream) {
  if (edge_dev == nullptr) {
    tmf::set_error("tmf_slater_modes_nested: edge_dev is required");
    return TMF_ERR_VALUE;
  }
  return modes_impl(C_dev, L, ldc, njobs, job_x, job_side, cutoff, r_sketch, v_off, V_dev, e_dev, info_dev, work_dev,
                    work_bytes, stream, true, edge_dev);
}

// The same for the re/im-interleaved real embedding (2L x 2L, cuts at 2x) of a complex Hermitian projector
// (complex Slater determinants, slater.py:1150-1180 keeps complex C): the per-bond eigenproblems of the complex
// blocks (slater.py:347, zheevd in the reference) run through the real kernels, every complex eigenvector showing
// up as a doubly degenerate real pair; pair_complex_kernel then picks the k complex modes (interleaved complex
// = the same memory as the real columns) and the edge vector is the embedding of the complex one.
// On return: info[0] = k complex modes, info[1] = f complex filled orbitals, e[0..k) their left eigenvalues.
extern "C" int tmf_slater_modes_nested_emb(const double *Cemb_dev, int L2, int ldc, int njobs,
                                           const int *job_x2, const int *job_side, double cutoff,
                                           int r_sketch, const int64_t *v_off, double *V_dev,
                                           double *e_dev, int *info_dev, double *edge_dev, void *work_dev,
                                           int64_t work_bytes, void *stream) {
  if (edge_dev == nullptr || (L2 & 1)) {
    tmf::set_error("tmf_slater_modes_nested_emb: edge_dev is required and the embedded size must be even");
    return TMF_ERR_VALUE;
  }
  return modes_impl(Cemb_dev, L2, ldc, njobs, job_x2, job_side, cutoff, r_sketch, v_off, V_dev, e_dev, info_dev,
                    work_dev, work_bytes, stream, true, edge_dev, true);
}
