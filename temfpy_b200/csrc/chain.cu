// Native driver of the Slater -> MPS chain: replaces the per-site Python loop of
// slater.C_to_MPS (slater.py:1216-1353) for a contiguous range of sites [site_lo, site_hi).
//
//   tmf_chain_modes    K3 for every needed (bond, side), one D2H of the (tiny) spectra, then on the
//                      host (threads over bonds / sites): best-first enumeration (K6/K7), sector
//                      tables, per-site planning;
//   tmf_chain_tensors  centre-bond pairing (K4), one upload of all plans, K8+K9 and K10 launches.
// Device buffers are owned by the caller (PyTorch); results (Schmidt values, charges, sector and
// block tables) stay in host vectors owned by the chain object until tmf_chain_destroy.
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <mutex>
#include <thread>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "cta.hpp"
#include "hostlogic.hpp"
#include "plan.hpp"

namespace {
struct StageTimer {   // TMF_DEBUG_TIMING=1 prints host-side stage times to stderr
  bool on;
  std::chrono::steady_clock::time_point t;
  StageTimer() : on(std::getenv("TMF_DEBUG_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
  void lap(const char *what) {
    if (!on) return;
    auto n = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[tmf timing] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};
}  // namespace

namespace tmf {
int gemm_grouped(const tmf_gemm_job *jobs, int njobs, void *desc_dev, void *stream);
int64_t gemm_desc_bytes(int njobs);
}  // namespace tmf

namespace tmf {
int64_t enum_workspace_bytes(int nb, int chi_max);
int enumerate_device(int nb, const double *const *e_ptr, const int *k, const int *filled_left, const TruncPar &tp,
                     std::vector<BondVectors *> &out, void *work_dev, int64_t work_bytes, void *stream,
                     unsigned char *stage, size_t stage_bytes, int n_threads, EnumResident *keep);
size_t enumerate_tables_stage_bytes(const EnumResident &r);
int enumerate_fetch_tables(const EnumResident &r, unsigned char *stage, void *stream);
void enumerate_unpack_tables(const EnumResident &r, const unsigned char *stage, const int *k, const int *filled_left,
                             std::vector<BondVectors *> &out, int n_threads);
}  // namespace tmf

extern "C" int64_t tmf_site_desc_bytes(int nsites);
extern "C" int64_t tmf_minor_desc_bytes(int nblocks);

struct ChainSide {
  int job = -1, n = 0, k = 0, f = 0;
  int64_t v_off = 0;
};
struct ChainBond {
  bool used = false;
  ChainSide side[2];  // TMF_SIDE_L, TMF_SIDE_R
  int k = 0, filled_left = 0;
  tmf::BondVectors bv;
};
struct ChainSite {
  int site = 0, mode = 0, bra_bond = 0, ket_bond = 0;
  int f_common = 0, df = 0;   // nested mode: filled orbitals shared by the two bonds, f_ket - f_bra
  // device-planned sites: the per-site arrays live in the caller's enumeration workspace
  const uint64_t *bra_masks_dev = nullptr, *ket_masks_dev = nullptr;
  const int *cols_dev = nullptr;
  const double *signs_dev = nullptr;
  bool rows_ready = false;    // row_p / row_alpha derived on the host (lazily) from the bra bond's charges
  tmf::SitePlan plan;
  int64_t o_off = 0, s_off = 0;
  std::vector<int64_t> block_off;
};

struct tmf_chain {
  int L = 0, oc = 0, nferm = 0, r_sketch = 64, site_lo = 0, site_hi = 0, n_threads = 0;
  tmf::TruncPar tp;
  std::vector<int> job_x, job_side;
  std::vector<int64_t> v_off;
  int64_t v_elems = 0;
  std::vector<ChainBond> bonds;
  std::vector<ChainSite> sites;
  std::vector<double> e_host;
  std::vector<int> info_host;
  int64_t o_elems = 0, s_elems = 0, out_elems = 0, plan_bytes = 0;
  int nblocks = 0, max_chi = 0;
  bool enumerated = false;
  bool nested = true;                // nested-projector site stage (no filled bases), see siteprep.cu
  bool peer_out = false;             // out_dev of the tensor stage is a peer window of another GPU (TMF_OPT_PEER_OUT)
  bool cplx = false;                 // complex Slater determinant: C_dev is the 2L x 2L real embedding (TMF_OPT_COMPLEX)
  std::vector<int> job_x2;           // cplx: cut positions in the embedded matrix (2 x)
  bool want_device_plan = true;      // plan the sites on the device when the enumeration ran there (nested mode)
  bool device_plan = false;          // ... and it did: bond tables + site plans are resident in the workspace
  bool tables_pending = false;       // bond tables still have to be unpacked from `tab_stage`
  tmf::EnumResident enum_res;
  std::vector<int> used_bonds;       // bond of every enumeration slot
  unsigned char *tab_stage = nullptr;  // pinned staging of the bond tables (asynchronous download)
  size_t tab_cap = 0;
  void *tab_event = nullptr;         // cudaEvent_t recorded after that download
  std::mutex tab_mu;
  const double *e_dev_ptr = nullptr; // device spectra of the last tmf_chain_modes_enqueue (read by the nested site kernel)
  std::vector<double> edge_host;     // nested: {|P_F e_edge|^2, rounding remainder of f} per job
  unsigned char *blob = nullptr;     // pinned staging of the per-site plan arrays (from the pool below)
  size_t blob_cap = 0;
  ~tmf_chain();
};

namespace {
// Process-wide pool of pinned host buffers: the plan blob (tens of MB per chain) is uploaded with one
// asynchronous copy; pageable memory would make that copy synchronous and ~5x slower, and pinning a
// fresh buffer per conversion would cost more than the copy.
struct PinnedPool {
  std::mutex mu;
  std::vector<std::pair<unsigned char *, size_t>> free_list;
  unsigned char *acquire(size_t bytes, size_t &cap) {
    {
      std::lock_guard<std::mutex> lk(mu);
      int best = -1;                       // best fit: big buffers stay available for big requests
      for (size_t i = 0; i < free_list.size(); ++i)
        if (free_list[i].second >= bytes && (best < 0 || free_list[i].second < free_list[best].second)) best = (int)i;
      if (best >= 0) {
        auto r = free_list[best];
        free_list.erase(free_list.begin() + best);
        cap = r.second;
        return r.first;
      }
    }
    cap = std::max<size_t>(bytes + bytes / 4, 1 << 20);
    unsigned char *p = nullptr;
#if defined(TMF_HOSTSIM)
    p = static_cast<unsigned char *>(std::malloc(cap));
#else
    if (cudaHostAlloc(reinterpret_cast<void **>(&p), cap, cudaHostAllocDefault) != cudaSuccess) p = nullptr;
#endif
    return p;
  }
  void release(unsigned char *p, size_t cap) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu);
    // (cudaFreeHost synchronises the device: keep enough buffers for several conversions in flight -- chains are
    //  destroyed by a background thread, a step can start before the previous one has returned its buffers)
    if (free_list.size() < 256) { free_list.emplace_back(p, cap); return; }
#if defined(TMF_HOSTSIM)
    std::free(p);
#else
    cudaFreeHost(p);
#endif
  }
};
PinnedPool g_pinned;
}  // namespace

tmf_chain::~tmf_chain() {
  g_pinned.release(blob, blob_cap);
  g_pinned.release(tab_stage, tab_cap);
#if !defined(TMF_HOSTSIM)
  if (tab_event) cudaEventDestroy((cudaEvent_t)tab_event);
#endif
}


namespace {

template <class F>
void parallel_for(int n, int n_threads, F f) {
  // exceptions keep their class through the pool: invalid_argument -> ValueError, others -> AssertionError
  try {
    tmf::pool_for(n, n_threads, [&](int i) { f(i); });
  } catch (const std::invalid_argument &e) {
    throw std::runtime_error(std::to_string(TMF_ERR_VALUE) + "|" + e.what());
  } catch (const std::runtime_error &e) {
    const std::string w = e.what();
    if (!w.empty() && w[0] == '-' && w.find('|') != std::string::npos) throw;
    throw std::runtime_error(std::to_string(TMF_ERR_ASSERT) + "|" + w);
  } catch (const std::exception &e) {
    throw std::runtime_error(std::to_string(TMF_ERR_ASSERT) + "|" + e.what());
  }
}

// one-sided Jacobi SVD of a small m x m matrix (column-major): M = U diag(s) V^T
void small_svd(std::vector<double> G, int m, std::vector<double> &U, std::vector<double> &V) {
  V.assign((size_t)m * m, 0.0);
  for (int i = 0; i < m; ++i) V[(size_t)i * m + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool any = false;
    for (int p = 0; p < m; ++p)
      for (int q = p + 1; q < m; ++q) {
        double a = 0, b = 0, c = 0;
        for (int r = 0; r < m; ++r) {
          a += G[p * m + r] * G[p * m + r];
          b += G[q * m + r] * G[q * m + r];
          c += G[p * m + r] * G[q * m + r];
        }
        if (c == 0.0 || std::fabs(c) <= 1e-15 * std::sqrt(a) * std::sqrt(b)) continue;
        any = true;
        double zeta = (b - a) / (2 * c);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
        double cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
        for (int r = 0; r < m; ++r) {
          double x = G[p * m + r], y = G[q * m + r];
          G[p * m + r] = cs * x - sn * y;
          G[q * m + r] = sn * x + cs * y;
          x = V[p * m + r]; y = V[q * m + r];
          V[p * m + r] = cs * x - sn * y;
          V[q * m + r] = sn * x + cs * y;
        }
      }
    if (!any) break;
  }
  U.assign((size_t)m * m, 0.0);
  for (int c = 0; c < m; ++c) {
    double s = 0;
    for (int r = 0; r < m; ++r) s += G[c * m + r] * G[c * m + r];
    s = std::sqrt(s);
    for (int r = 0; r < m; ++r) U[c * m + r] = (s > 0) ? G[c * m + r] / s : (r == c ? 1.0 : 0.0);
  }
}

int fail(int code, const std::string &msg) {
  tmf::set_error(msg);
  return code;
}
int fail_from(const std::exception &e) {
  std::string w = e.what();
  auto bar = w.find('|');
  if (bar != std::string::npos && (w[0] == '-' )) {
    int code = std::atoi(w.substr(0, bar).c_str());
    tmf::set_error(w.substr(bar + 1));
    return code;
  }
  tmf::set_error(w);
  return TMF_ERR_RUNTIME;
}

}  // namespace

// V slots of every job: n x n ([entangled | filled basis]) in the legacy layout, n x (Ritz columns + 1)
// ([entangled | edge vector]) in the nested one
static void chain_layout_slots(tmf_chain *c) {
  c->v_elems = 0;
  c->job_x2.resize(c->job_x.size());
  for (size_t j = 0; j < c->job_x.size(); ++j) c->job_x2[j] = 2 * c->job_x[j];
  for (size_t j = 0; j < c->job_x.size(); ++j) {
    const int bond = c->job_x[j], side = c->job_side[j];
    ChainSide &s = c->bonds[bond].side[side];
    // (complex: the slot holds the real columns of the embedded 2n-row block; afterwards the same memory read as
    //  interleaved complex columns of n rows)
    const int es = c->cplx ? 2 : 1;
    const int64_t cols = tmf_slater_modes_slot_cols(es * c->L, es * bond, side, c->r_sketch, c->nested ? 1 : 0);
    s.v_off = c->v_elems;
    c->v_off[j] = s.v_off;
    c->v_elems += (int64_t)es * s.n * cols + 32;  // + slack keeps every matrix 256-byte aligned
    c->v_elems = (c->v_elems + 31) & ~int64_t(31);
  }
}

extern "C" {

tmf_chain *tmf_chain_create(int L, int ortho_center, int n_fermion, int chi_max, double svd_min,
                            double degeneracy_tol, const int *sectors, int n_sectors, int r_sketch,
                            int site_lo, int site_hi, int n_threads) {
  if (L <= 0 || ortho_center < 0 || ortho_center > L || site_lo < 0 || site_hi > L || site_lo >= site_hi) {
    tmf::set_error("tmf_chain_create: bad geometry");
    return nullptr;
  }
  tmf_chain *c = new tmf_chain();
  c->L = L; c->oc = ortho_center; c->nferm = n_fermion; c->r_sketch = r_sketch;
  c->site_lo = site_lo; c->site_hi = site_hi; c->n_threads = n_threads;
  c->tp.chi_max = chi_max; c->tp.svd_min = svd_min; c->tp.degeneracy_tol = degeneracy_tol;
  if (sectors && n_sectors >= 0) {
    c->tp.filter = true;
    c->tp.sectors.assign(sectors, sectors + n_sectors);
  }
  c->bonds.resize(L + 1);
  c->nested = std::getenv("TMF_LEGACY_FILLED") == nullptr;
  auto need = [&](int bond, int side) {
    ChainBond &b = c->bonds[bond];
    b.used = true;
    if (b.side[side].job >= 0) return;
    ChainSide &s = b.side[side];
    s.job = (int)c->job_x.size();
    s.n = (side == TMF_SIDE_L) ? bond : L - bond;
    c->job_x.push_back(bond);
    c->job_side.push_back(side);
    c->v_off.push_back(0);
  };
  for (int i = site_lo; i < site_hi; ++i) {
    const int side = (i >= ortho_center) ? TMF_SIDE_R : TMF_SIDE_L;
    need(i, side);
    need(i + 1, side);
    if (i == ortho_center || i + 1 == ortho_center) {  // centre bond: both sides + pairing
      need(ortho_center, TMF_SIDE_L);
      need(ortho_center, TMF_SIDE_R);
    }
  }
  chain_layout_slots(c);
  return c;
}

int tmf_chain_set_option(tmf_chain *c, int option, int value) {
  if (option == TMF_OPT_SNAP) { c->tp.snap = value != 0; return TMF_OK; }
  if (option == TMF_OPT_NESTED) {
    c->nested = value != 0;
    chain_layout_slots(c);
    return TMF_OK;
  }
  if (option == TMF_OPT_DEVICE_PLAN) { c->want_device_plan = value != 0; return TMF_OK; }
  if (option == TMF_OPT_PEER_OUT) { c->peer_out = value != 0; return TMF_OK; }
  if (option == TMF_OPT_COMPLEX) {
    c->cplx = value != 0;
    if (c->cplx) c->nested = true;      // the complex kernels exist in the nested form only
    chain_layout_slots(c);
    return TMF_OK;
  }
  tmf::set_error("tmf_chain_set_option: unknown option");
  return TMF_ERR_VALUE;
}

void tmf_chain_destroy(tmf_chain *c) { delete c; }

int tmf_chain_modes_sizes(tmf_chain *c, int64_t *q) {
  q[0] = (int64_t)c->job_x.size();
  q[1] = c->v_elems;
  q[2] = c->cplx ? tmf_slater_modes_workspace(2 * c->L, (int)c->job_x.size(), c->job_x2.data(), c->job_side.data(), c->r_sketch)
                 : tmf_slater_modes_workspace(c->L, (int)c->job_x.size(), c->job_x.data(), c->job_side.data(), c->r_sketch);
  q[3] = (int64_t)c->job_x.size() * (TMF_MAX_MODES + 2);   // doubles of e_dev (spectra + edge data)
  return TMF_OK;
}

// Enqueues every mode-extraction kernel of the shard on `stream` and returns without waiting
// (tmf_chain_modes_finish brings the spectra to the host).  The split lets a pipeline driver record
// an event between the two and order the mode stages of consecutive chunks on the device.
int tmf_chain_modes_enqueue(tmf_chain *c, const double *C_dev, int ldc, double *V_dev, double *e_dev,
                            int *info_dev, void *work_dev, int64_t work_bytes, void *stream) {
  const int nj = (int)c->job_x.size();
  const double cutoff = c->tp.svd_min * c->tp.svd_min;  // slater.py:318
  c->e_dev_ptr = e_dev;
  if (c->cplx)
    return tmf_slater_modes_nested_emb(C_dev, 2 * c->L, ldc, nj, c->job_x2.data(), c->job_side.data(), cutoff,
                                       c->r_sketch, c->v_off.data(), V_dev, e_dev, info_dev,
                                       e_dev + (size_t)nj * TMF_MAX_MODES, work_dev, work_bytes, stream);
  if (c->nested)
    return tmf_slater_modes_nested(C_dev, c->L, ldc, nj, c->job_x.data(), c->job_side.data(), cutoff,
                                   c->r_sketch, c->v_off.data(), V_dev, e_dev, info_dev,
                                   e_dev + (size_t)nj * TMF_MAX_MODES, work_dev, work_bytes, stream);
  return tmf_slater_modes_batched(C_dev, c->L, ldc, nj, c->job_x.data(), c->job_side.data(), cutoff,
                                  c->r_sketch, c->v_off.data(), V_dev, e_dev, info_dev, work_dev,
                                  work_bytes, stream);
}

int tmf_chain_modes_finish(tmf_chain *c, const double *e_dev, const int *info_dev, void *stream) {
  const int nj = (int)c->job_x.size();
  c->e_host.resize((size_t)nj * TMF_MAX_MODES);
  c->info_host.resize((size_t)nj * 4);
  // through pinned staging: a device -> pageable copy is executed synchronously inside the driver (it waits
  // there for the whole mode stage of this chunk), which stalled the kernel launches of the other pipeline
  // threads for milliseconds; cudaStreamSynchronize does not
  const size_t e_bytes = sizeof(double) * (TMF_MAX_MODES + (c->nested ? 2 : 0)) * (size_t)nj, i_bytes = sizeof(int) * 4 * (size_t)nj;
  c->edge_host.assign(c->nested ? 2 * (size_t)nj : 0, 0.0);
  const size_t need = e_bytes + i_bytes + 512;
  if (c->blob_cap < need) {
    g_pinned.release(c->blob, c->blob_cap);
    c->blob = g_pinned.acquire(need, c->blob_cap);
    if (!c->blob) { c->blob_cap = 0; return fail(TMF_ERR_RUNTIME, "cannot allocate pinned staging memory"); }
  }
  unsigned char *st_e = c->blob, *st_i = c->blob + ((e_bytes + 255) & ~size_t(255));
  int rc = tmf::copy_d2h_async(st_i, info_dev, i_bytes, stream);
  if (rc) return rc;
  rc = tmf::copy_d2h_sync(st_e, e_dev, e_bytes, stream);
  if (rc) return rc;
  std::memcpy(c->e_host.data(), st_e, sizeof(double) * TMF_MAX_MODES * (size_t)nj);
  if (c->nested) std::memcpy(c->edge_host.data(), st_e + sizeof(double) * TMF_MAX_MODES * (size_t)nj, sizeof(double) * 2 * (size_t)nj);
  std::memcpy(c->info_host.data(), st_i, i_bytes);
  for (int j = 0; j < nj; ++j) {
    const int st = c->info_host[4 * j + 2];
    if (st & 2) return fail(TMF_ERR_VALUE, "more than 64 entangled modes on one bond");
    if (st & 4) return fail(TMF_ERR_RUNTIME, "complex modes: the real eigenvector pairs of the embedded matrix could not be paired");
    if (st & 1)
      return fail(TMF_ERR_VALUE,
                  "range sketch too narrow for this entanglement spectrum: rerun with a larger r_sketch");
  }
  return TMF_OK;
}

int tmf_chain_modes(tmf_chain *c, const double *C_dev, int ldc, double *V_dev, double *e_dev,
                    int *info_dev, void *work_dev, int64_t work_bytes, void *stream) {
  int rc = tmf_chain_modes_enqueue(c, C_dev, ldc, V_dev, e_dev, info_dev, work_dev, work_bytes, stream);
  if (rc) return rc;
  return tmf_chain_modes_finish(c, e_dev, info_dev, stream);
}

// device buffers of the site planner (plan.cu), carved from the enumeration workspace after the bond tables
static int64_t plan_site_bytes(int64_t cap) {
  return tmf::align256(16 * cap) + tmf::align256(8 * cap) + tmf::align256(4 * 2 * tmf::PLAN_MAX_ORB) +
         tmf::align256(8 * 2 * tmf::PLAN_MAX_ORB);
}
static int64_t plan_workspace_bytes(int ns, int64_t cap) {
  return (int64_t)ns * plan_site_bytes(cap) + tmf::align256((int64_t)ns * sizeof(tmf::PlanJob)) +
         tmf::align256((int64_t)ns * 4 * (tmf::PLAN_HDR_INTS + 6 * tmf::PLAN_MAX_BLOCKS)) + 4096;
}

// bra rows of a site in the reference order: [p = 0 | p = 1], stably sorted by the pipe charge (slater.py:1053-1058)
static void rows_from_charges(int mode, int chi_b, const int *charge_b, std::vector<int> &row_p, std::vector<int> &row_alpha) {
  const int n_rows = 2 * chi_b;
  row_p.resize(n_rows); row_alpha.resize(n_rows);
  if (chi_b == 0) return;
  int qmin = 1 << 30, qmax = -(1 << 30);
  for (int a = 0; a < chi_b; ++a) { qmin = std::min(qmin, charge_b[a] - 1); qmax = std::max(qmax, charge_b[a] + 1); }
  std::vector<int> start((size_t)(qmax - qmin + 2), 0);
  auto rq = [&](int r) { const int p = r / chi_b, a = r % chi_b; return charge_b[a] + (mode == 0 ? p : -p); };
  for (int r = 0; r < n_rows; ++r) ++start[rq(r) - qmin + 1];
  for (size_t q = 1; q < start.size(); ++q) start[q] += start[q - 1];
  for (int r = 0; r < n_rows; ++r) {
    const int pos = start[rq(r) - qmin]++;
    row_p[pos] = r / chi_b; row_alpha[pos] = r % chi_b;
  }
}

// enumeration (device kernel when a workspace is given, host otherwise) + planning (host).  Requires
// tmf_chain_modes to have completed.
static int chain_enumerate_impl(tmf_chain *c, void *work_dev, int64_t work_bytes, void *stream);
int tmf_chain_enumerate(tmf_chain *c) { return chain_enumerate_impl(c, nullptr, 0, nullptr); }
int64_t tmf_chain_enum_workspace(tmf_chain *c) {
  int nb = 0;
  for (int b = 0; b <= c->L; ++b)
    if (c->bonds[b].used) ++nb;
  int64_t bytes = tmf::enum_workspace_bytes(nb, c->tp.chi_max);
  if (c->tp.chi_max >= 0) bytes += plan_workspace_bytes(c->site_hi - c->site_lo, c->tp.chi_max + 2);
  return bytes;
}
int tmf_chain_enumerate_dev(tmf_chain *c, void *work_dev, int64_t work_bytes, void *stream) {
  return chain_enumerate_impl(c, work_dev, work_bytes, stream);
}
// Site plans on the device (plan.cu) from the resident bond tables; only the headers and block tables
// (~1.7 KB per site) come back to the host, which needs them for the output offsets and the descriptors.
static int chain_plan_device(tmf_chain *c, void *work_dev, int64_t work_bytes, void *stream) {
  const tmf::EnumResident &er = c->enum_res;
  const int ns = (int)c->sites.size();
  const int64_t cap = er.cap;
  std::vector<int> slot(c->L + 1, -1);
  for (size_t u = 0; u < c->used_bonds.size(); ++u) slot[c->used_bonds[u]] = (int)u;
  tmf::Arena ar(work_dev, work_bytes);
  ar.take<unsigned char>(tmf::enum_workspace_bytes(er.nb, c->tp.chi_max));
  tmf::PlanJob *jobs_dev = ar.take<tmf::PlanJob>(ns);
  const int HB = tmf::PLAN_HDR_INTS + 6 * tmf::PLAN_MAX_BLOCKS;
  int *hb_dev = ar.take<int>((int64_t)ns * HB);
  std::vector<tmf::PlanJob> jobs(ns);
  for (int u = 0; u < ns; ++u) {
    ChainSite &s = c->sites[u];
    const int side = s.mode == 1 ? TMF_SIDE_R : TMF_SIDE_L;
    const ChainBond &bb = c->bonds[s.bra_bond], &kb = c->bonds[s.ket_bond];
    const ChainSide &bs = bb.side[side], &ks = kb.side[side];
    s.df = ks.f - bs.f;
    s.f_common = bs.f;
    if (s.df < 0 || s.df > 1)
      throw std::invalid_argument("nested site stage: filled-orbital counts of neighbouring bonds differ by " +
                                  std::to_string(s.df) + " (threshold noise); rerun with the legacy filled bases");
    if (s.df == 1 && !(c->edge_host[2 * (size_t)ks.job] > 1e-24))
      throw std::invalid_argument("nested site stage: edge vector of the filled space vanishes; rerun with the "
                                  "legacy filled bases");
    uint64_t *bm = ar.take<uint64_t>(2 * cap), *km = ar.take<uint64_t>(cap);
    int *cols = ar.take<int>(2 * tmf::PLAN_MAX_ORB);
    double *signs = ar.take<double>(2 * tmf::PLAN_MAX_ORB);
    s.bra_masks_dev = bm; s.ket_masks_dev = km; s.cols_dev = cols; s.signs_dev = signs;
    s.rows_ready = false;
    const int sb = slot[s.bra_bond], sk = slot[s.ket_bond];
    tmf::PlanJob &j = jobs[u];
    std::memset(&j, 0, sizeof(j));
    j.masks_b = er.masks_dev + (int64_t)sb * cap; j.masks_k = er.masks_dev + (int64_t)sk * cap;
    j.charge_b = er.charge_dev + (int64_t)sb * cap; j.charge_k = er.charge_dev + (int64_t)sk * cap;
    j.head_b = er.head_dev + (int64_t)sb * er.hw; j.head_k = er.head_dev + (int64_t)sk * er.hw;
    j.bra_masks = bm; j.ket_masks = km; j.cols = cols; j.signs = signs;
    j.hdr = hb_dev + (int64_t)u * HB; j.blocks = j.hdr + tmf::PLAN_HDR_INTS;
    j.mode = s.mode; j.k_bra = bb.k; j.k_ket = kb.k; j.df = s.df;
    j.n_bra = bs.n; j.n_ket = ks.n; j.f_bra = bs.f; j.f_ket = ks.f;
  }
  if (!ar.ok()) return fail(TMF_ERR_VALUE, "enumeration workspace too small for the device site plans");
  int rc = tmf::copy_h2d(jobs_dev, jobs.data(), sizeof(tmf::PlanJob) * (size_t)ns, stream);
  if (rc) return rc;
  rc = tmf::plan_sites_device(jobs_dev, ns, stream);
  if (rc) return rc;
  // headers + block tables through the pinned staging (the enumeration's job staging is dead by now)
  const size_t hb_bytes = sizeof(int) * (size_t)ns * HB;
  if (c->blob_cap < hb_bytes + 256) {
    g_pinned.release(c->blob, c->blob_cap);
    c->blob = g_pinned.acquire(hb_bytes + 256, c->blob_cap);
    if (!c->blob) { c->blob_cap = 0; return fail(TMF_ERR_RUNTIME, "cannot allocate pinned staging memory"); }
  }
  rc = tmf::copy_d2h_sync(c->blob, hb_dev, hb_bytes, stream);
  if (rc) return rc;
  const int *hb = reinterpret_cast<const int *>(c->blob);
  for (int u = 0; u < ns; ++u) {
    ChainSite &s = c->sites[u];
    const int *h = hb + (size_t)u * HB;
    if (h[18] == 2) throw std::invalid_argument("sometimes matrix larger than 64 rows/cols is not supported");
    if (h[18] != 0) throw std::runtime_error("particle numbers of bra and ket block differ (slater.py:855)");
    std::memcpy(&s.plan.h, h, sizeof(tmf_site_plan));
    s.plan.blocks.assign(h + tmf::PLAN_HDR_INTS, h + tmf::PLAN_HDR_INTS + 6 * s.plan.h.n_blocks);
    s.plan.bra_cols.clear(); s.plan.ket_cols.clear(); s.plan.bra_masks.clear(); s.plan.ket_masks.clear();
    s.plan.row_p.clear(); s.plan.row_alpha.clear();
  }
  c->tables_pending = true;
  return TMF_OK;
}

// Brings the bond tables of a device-planned chain to the host (once): waits for the asynchronous download
// enqueued by tmf_chain_tensors (or performs it now) and unpacks it into the per-bond vectors.
static int chain_ensure_tables(tmf_chain *c) {
  if (!c->device_plan || !c->tables_pending) return TMF_OK;
  std::lock_guard<std::mutex> lk(c->tab_mu);
  if (!c->tables_pending) return TMF_OK;
  const tmf::EnumResident &er = c->enum_res;
#if !defined(TMF_HOSTSIM)
  if (c->tab_event) {
    int rc = tmf::check_cuda(cudaEventSynchronize((cudaEvent_t)c->tab_event), "cudaEventSynchronize");
    if (rc) return rc;
  } else
#endif
  {
    const size_t need = tmf::enumerate_tables_stage_bytes(er);
    if (c->tab_cap < need) {
      g_pinned.release(c->tab_stage, c->tab_cap);
      c->tab_stage = g_pinned.acquire(need, c->tab_cap);
      if (!c->tab_stage) { c->tab_cap = 0; return fail(TMF_ERR_RUNTIME, "cannot allocate pinned staging memory"); }
    }
    int rc = tmf::enumerate_fetch_tables(er, c->tab_stage, nullptr);
    if (rc) return rc;
    rc = tmf::stream_sync(nullptr);
    if (rc) return rc;
  }
  const int nbu = (int)c->used_bonds.size();
  std::vector<int> kk(nbu), fl(nbu);
  std::vector<tmf::BondVectors *> outp(nbu);
  for (int u = 0; u < nbu; ++u) {
    ChainBond &B = c->bonds[c->used_bonds[u]];
    kk[u] = B.k; fl[u] = B.filled_left; outp[u] = &B.bv;
  }
  tmf::enumerate_unpack_tables(er, c->tab_stage, kk.data(), fl.data(), outp, c->n_threads > 0 ? c->n_threads : 4);
  c->tables_pending = false;
  return TMF_OK;
}
static void chain_site_rows(tmf_chain *c, ChainSite &s) {
  if (!c->device_plan || s.rows_ready) return;
  const ChainBond &bb = c->bonds[s.bra_bond];
  rows_from_charges(s.mode, (int)bb.bv.charge.size(), bb.bv.charge.data(), s.plan.row_p, s.plan.row_alpha);
  s.rows_ready = true;
}

static int chain_enumerate_impl(tmf_chain *c, void *work_dev, int64_t work_bytes, void *stream) {
  try {
    StageTimer tm;
    std::vector<int> used;
    for (int b = 0; b <= c->L; ++b)
      if (c->bonds[b].used) used.push_back(b);
    for (int b : used) {
      ChainBond &B = c->bonds[b];
      for (int s = 0; s < 2; ++s)
        if (B.side[s].job >= 0) {
          B.side[s].k = c->info_host[4 * B.side[s].job];
          B.side[s].f = c->info_host[4 * B.side[s].job + 1];
        }
      const ChainSide &Ls = B.side[TMF_SIDE_L], &Rs = B.side[TMF_SIDE_R];
      if (Ls.job >= 0 && Rs.job >= 0 && Ls.k != Rs.k)
        throw std::runtime_error("-2|number of entangled modes differs between the two sides (slater.py:394)");
      B.k = (Ls.job >= 0) ? Ls.k : Rs.k;
      // slater.py:145-174: filled-left count, inferred from n_fermion when only vR is known
      B.filled_left = (Ls.job >= 0) ? Ls.f : c->nferm - B.k - Rs.f;
    }
    if (work_dev != nullptr) {
      const int nbu = (int)used.size();
      std::vector<const double *> ep(nbu);
      std::vector<int> kk(nbu), fl(nbu);
      std::vector<tmf::BondVectors *> outp(nbu);
      for (int u = 0; u < nbu; ++u) {
        ChainBond &B = c->bonds[used[u]];
        const int job = (B.side[TMF_SIDE_L].job >= 0) ? B.side[TMF_SIDE_L].job : B.side[TMF_SIDE_R].job;
        ep[u] = c->e_host.data() + (size_t)job * TMF_MAX_MODES;
        kk[u] = B.k; fl[u] = B.filled_left; outp[u] = &B.bv;
      }
      // pinned staging for the job upload and the table download (shared with the plan blob later)
      const size_t need = (size_t)nbu * (2048 + (size_t)(std::max(c->tp.chi_max, 0) + 2) * 20 + 1024) + 4096;
      if (c->blob_cap < need) {
        g_pinned.release(c->blob, c->blob_cap);
        c->blob = g_pinned.acquire(need, c->blob_cap);
        if (!c->blob) c->blob_cap = 0;
      }
      c->device_plan = false;
      c->used_bonds = used;
      const bool try_resident = c->nested && c->want_device_plan && c->tp.chi_max >= 0;
      int rc = tmf::enumerate_device(nbu, ep.data(), kk.data(), fl.data(), c->tp, outp, work_dev, work_bytes, stream,
                                     c->blob, c->blob_cap, c->n_threads > 0 ? c->n_threads : 8,
                                     try_resident ? &c->enum_res : nullptr);
      if (rc) return rc;
      c->device_plan = try_resident && c->enum_res.resident;
    } else {
    c->device_plan = false;
    parallel_for((int)used.size(), c->n_threads, [&](int u) {
      ChainBond &B = c->bonds[used[u]];
      const int job = (B.side[TMF_SIDE_L].job >= 0) ? B.side[TMF_SIDE_L].job : B.side[TMF_SIDE_R].job;
      tmf::bond_vectors(c->e_host.data() + (size_t)job * TMF_MAX_MODES, B.k, B.filled_left, c->tp, B.bv);
    });
    }
    tm.lap("enumerate: bond vectors");
    c->sites.clear();
    for (int i = c->site_lo; i < c->site_hi; ++i) {
      ChainSite s;
      s.site = i;
      if (i >= c->oc) { s.mode = 1; s.bra_bond = i + 1; s.ket_bond = i; }   // slater.py:1301-1310
      else { s.mode = 0; s.bra_bond = i; s.ket_bond = i + 1; }               // slater.py:1326-1335
      c->sites.push_back(std::move(s));
    }
    if (c->device_plan) {
      int rc = chain_plan_device(c, work_dev, work_bytes, stream);
      if (rc) return rc;
      tm.lap("enumerate: site plans (device)");
    } else
    parallel_for((int)c->sites.size(), c->n_threads, [&](int u) {
      ChainSite &s = c->sites[u];
      const int side = s.mode == 1 ? TMF_SIDE_R : TMF_SIDE_L;
      const ChainBond &bb = c->bonds[s.bra_bond], &kb = c->bonds[s.ket_bond];
      const ChainSide &bs = bb.side[side], &ks = kb.side[side];
      if (!c->nested) {
        tmf::site_plan(s.mode, bs.n, ks.n, bb.k, bs.f, c->nferm, (int)bb.bv.masks.size(),
                       bb.bv.masks.data(), bb.bv.charge.data(), kb.k, ks.f, c->nferm,
                       (int)kb.bv.masks.size(), kb.bv.masks.data(), kb.bv.charge.data(), s.plan);
        return;
      }
      // nested: the filled orbitals the two bonds share are eliminated in closed form (siteprep.cu); what is
      // left of them for the plan is the ket's edge vector (stored-V column k_ket) when its count grows
      s.df = ks.f - bs.f;
      s.f_common = bs.f;
      if (s.df < 0 || s.df > 1)
        throw std::invalid_argument("nested site stage: filled-orbital counts of neighbouring bonds differ by " +
                                    std::to_string(s.df) + " (threshold noise); rerun with the legacy filled bases");
      if (s.df == 1 && !(c->edge_host[2 * (size_t)ks.job] > 1e-24))
        throw std::invalid_argument("nested site stage: edge vector of the filled space vanishes; rerun with the "
                                    "legacy filled bases");
      tmf::site_plan(s.mode, bs.n, ks.n, bb.k, 0, c->nferm, (int)bb.bv.masks.size(),
                     bb.bv.masks.data(), bb.bv.charge.data(), kb.k, s.df, c->nferm,
                     (int)kb.bv.masks.size(), kb.bv.masks.data(), kb.bv.charge.data(), s.plan);
      s.plan.h.f_bra = bs.f;
      s.plan.h.f_ket = ks.f;
    });
    tm.lap("enumerate: site plans");
    // offsets
    c->o_elems = c->s_elems = c->out_elems = 0;
    c->nblocks = 0;
    c->max_chi = 0;
    int64_t plan = 0;
    for (ChainSite &s : c->sites) {
      const tmf_site_plan &h = s.plan.h;
      const int rows = h.ka_bra + (h.s_bra - (h.ka_bra - h.k_always)), cols = h.ka_ket + (h.s_ket - (h.ka_ket - h.k_always));
      s.o_off = c->o_elems;
      c->o_elems += ((int64_t)rows * cols + 31) & ~int64_t(31);
      s.s_off = c->s_elems;
      c->s_elems += ((int64_t)h.s_bra * h.s_ket + 31) & ~int64_t(31);
      s.block_off.clear();
      for (int b = 0; b < h.n_blocks; ++b) {
        s.block_off.push_back(c->out_elems);
        c->out_elems += (int64_t)s.plan.blocks[6 * b + 1] * s.plan.blocks[6 * b + 3];
      }
      c->nblocks += h.n_blocks;
      c->max_chi = std::max(c->max_chi, std::max(h.chi_bra, h.chi_ket));
      if (!c->device_plan)
        plan += tmf::align256(4 * (int64_t)rows) + tmf::align256(4 * (int64_t)cols) +
                tmf::align256(8 * (int64_t)rows) + tmf::align256(8 * (int64_t)cols) +
                tmf::align256(8 * (int64_t)h.n_rows) + tmf::align256(8 * (int64_t)h.chi_ket);
    }
    const int kc = c->bonds[c->oc].used ? c->bonds[c->oc].k : 0;
    const int64_t pair = (c->cplx ? tmf_slater_pair_bond_c_workspace(c->L, kc) : tmf_slater_pair_bond_workspace(c->L, kc)) + 512;
    c->plan_bytes = plan + tmf_site_desc_bytes((int)c->sites.size()) + tmf_minor_desc_bytes(c->nblocks) +
                    pair + 4096;
    c->enumerated = true;
    return TMF_OK;
  } catch (const std::exception &e) {
    return fail_from(e);
  }
}

// q = {plan_bytes, o_elems, s_elems, n_sites, n_blocks, out_elems, max_chi}
int tmf_chain_tensor_sizes(tmf_chain *c, int64_t *q) {
  if (!c->enumerated) return fail(TMF_ERR_VALUE, "tmf_chain_enumerate has not run");
  q[0] = c->plan_bytes; q[1] = c->o_elems; q[2] = c->s_elems; q[3] = (int64_t)c->sites.size();
  q[4] = c->nblocks; q[5] = c->out_elems; q[6] = c->max_chi;
  q[7] = (c->nested ? 1 : 0) | (c->device_plan ? 2 : 0);   // which site stage / planner this chain uses
  return TMF_OK;
}

int64_t tmf_slater_pair_bond_workspace(int L, int k) {
  return tmf::align256(8 * (int64_t)L * (k + 1)) * 3 + tmf::align256(8 * (int64_t)k * k) * 3 +
         tmf::gemm_desc_bytes(4) + 4096;
}

// K4: pairing of the left and right entangled modes of one bond (utils.py:19-96 as called from
// slater.py:407, and the sign flips of :410).  Rotates the first k columns of VL (x rows) and VR
// (L - x rows) in place.  Synchronises once (k x k matrix to the host for the small SVDs).
int tmf_slater_pair_bond(const double *C_dev, int ldc, int L, int x, int k, const double *e_host,
                         double degeneracy_tol, double *VL, double *VR, void *work_dev, int64_t work_bytes,
                         void *stream) {
  if (k <= 0) return TMF_OK;
  if (x <= 0 || x >= L) return fail(TMF_ERR_VALUE, "tmf_slater_pair_bond: bond has an empty side");
  tmf::Arena ar(work_dev, work_bytes);
  const int nL = x, nR = L - x;
  double *T1 = ar.take<double>((int64_t)nL * k), *tmpL = ar.take<double>((int64_t)nL * k);
  double *tmpR = ar.take<double>((int64_t)nR * k);
  double *M = ar.take<double>((int64_t)k * k), *RotL = ar.take<double>((int64_t)k * k);
  double *RotR = ar.take<double>((int64_t)k * k);
  void *desc = ar.take<unsigned char>(tmf::gemm_desc_bytes(4));
  if (!ar.ok()) return fail(TMF_ERR_VALUE, "workspace too small (pairing)");
  auto mk = [](const double *A, int lda, int tA, const double *Bm, int ldb, double *Cm, int ldc_, int Mm,
               int Nn, int Kk) {
    tmf_gemm_job j;
    std::memset(&j, 0, sizeof(j));
    j.A = A; j.B = Bm; j.C = Cm; j.lda = lda; j.ldb = ldb; j.ldc = ldc_;
    j.M = Mm; j.N = Nn; j.K = Kk; j.transA = tA; j.transB = 0; j.alpha = 1.0; j.beta = 0.0;
    return j;
  };
  // T1 = C_LR VR ;  M = VL^T T1   (utils.py:87-89)
  tmf_gemm_job j1 = mk(C_dev + (int64_t)x * ldc, ldc, 0, VR, nR, T1, nL, nL, k, nR);
  int rc = tmf::gemm_grouped(&j1, 1, desc, stream);
  if (rc) return rc;
  tmf_gemm_job j2 = mk(VL, nL, 1, T1, nL, M, k, k, k, nL);
  rc = tmf::gemm_grouped(&j2, 1, desc, stream);
  if (rc) return rc;
  std::vector<double> Mh((size_t)k * k), RL((size_t)k * k, 0.0), RR((size_t)k * k, 0.0);
  rc = tmf::copy_d2h_sync(Mh.data(), M, sizeof(double) * Mh.size(), stream);
  if (rc) return rc;
  const double *e = e_host;
  int a = 0;
  while (a < k) {  // groups of (nearly) degenerate eigenvalues (utils.py:71-78)
    int b = a + 1;
    while (b < k && !(std::fabs(e[b] - e[b - 1]) > degeneracy_tol)) ++b;
    const int m = b - a;
    std::vector<double> G((size_t)m * m), U, V;
    for (int cc = 0; cc < m; ++cc)
      for (int r = 0; r < m; ++r) G[(size_t)cc * m + r] = Mh[(size_t)(a + cc) * k + a + r];
    small_svd(G, m, U, V);
    for (int cc = 0; cc < m; ++cc)
      for (int r = 0; r < m; ++r) {
        RL[(size_t)(a + cc) * k + a + r] = U[(size_t)cc * m + r];
        RR[(size_t)(a + cc) * k + a + r] = V[(size_t)cc * m + r];
      }
    a = b;
  }
  // anticommutation signs (slater.py:410): reference column j of vRE is mode k-1-j; odd j flips
  for (int i = 0; i < k; ++i)
    if ((k - 1 - i) & 1)
      for (int r = 0; r < k; ++r) RR[(size_t)i * k + r] = -RR[(size_t)i * k + r];
  rc = tmf::copy_h2d(RotL, RL.data(), sizeof(double) * RL.size(), stream);
  if (rc) return rc;
  rc = tmf::copy_h2d(RotR, RR.data(), sizeof(double) * RR.size(), stream);
  if (rc) return rc;
  tmf_gemm_job j3[2] = {mk(VL, nL, 0, RotL, k, tmpL, nL, nL, k, k), mk(VR, nR, 0, RotR, k, tmpR, nR, nR, k, k)};
  rc = tmf::gemm_grouped(j3, 2, desc, stream);
  if (rc) return rc;
#if defined(TMF_HOSTSIM)
  std::memcpy(VL, tmpL, sizeof(double) * (size_t)nL * k);
  std::memcpy(VR, tmpR, sizeof(double) * (size_t)nR * k);
#else
  rc = tmf::check_cuda(cudaMemcpyAsync(VL, tmpL, sizeof(double) * (size_t)nL * k, cudaMemcpyDeviceToDevice,
                                       (cudaStream_t)stream), "pairing copy");
  if (rc) return rc;
  rc = tmf::check_cuda(cudaMemcpyAsync(VR, tmpR, sizeof(double) * (size_t)nR * k, cudaMemcpyDeviceToDevice,
                                       (cudaStream_t)stream), "pairing copy");
  if (rc) return rc;
#endif
  return TMF_OK;
}

static int centre_pairing(tmf_chain *c, const double *C_dev, int ldc, double *V_dev, tmf::Arena &ar,
                          void *stream) {
  ChainBond &B = c->bonds[c->oc];
  if (!B.used || B.side[0].job < 0 || B.side[1].job < 0 || B.k == 0) return TMF_OK;
  if (c->cplx) {
    const int64_t wbc = tmf_slater_pair_bond_c_workspace(c->L, B.k);
    void *workc = ar.take<unsigned char>(wbc);
    if (!ar.ok()) return fail(TMF_ERR_VALUE, "plan workspace too small (pairing)");
    return tmf_slater_pair_bond_c(C_dev, ldc, c->L, c->oc, B.k,
                                  c->e_host.data() + (size_t)B.side[TMF_SIDE_L].job * TMF_MAX_MODES,
                                  c->tp.degeneracy_tol, V_dev + B.side[TMF_SIDE_L].v_off,
                                  V_dev + B.side[TMF_SIDE_R].v_off, workc, wbc, stream);
  }
  const int64_t wb = tmf_slater_pair_bond_workspace(c->L, B.k);
  void *work = ar.take<unsigned char>(wb);
  if (!ar.ok()) return fail(TMF_ERR_VALUE, "plan workspace too small (pairing)");
  return tmf_slater_pair_bond(C_dev, ldc, c->L, c->oc, B.k,
                              c->e_host.data() + (size_t)B.side[TMF_SIDE_L].job * TMF_MAX_MODES,
                              c->tp.degeneracy_tol, V_dev + B.side[TMF_SIDE_L].v_off,
                              V_dev + B.side[TMF_SIDE_R].v_off, work, wb, stream);
}

// Tensor stage of a device-planned chain: the per-site arrays are already in device memory; the host only
// writes the kernel descriptors (pointers + sizes) and, after the last kernel, enqueues the download of the
// bond tables for the result accessors.
static int chain_tensors_device_plan(tmf_chain *c, const double *C_dev, int ldc, double *V_dev, tmf::Arena &ar,
                                     double *O_dev, double *S_dev, double *det_dev, double *out_dev, void *stream) {
  const int ns = (int)c->sites.size();
  std::vector<tmf_site_job> sj(ns);
  std::vector<tmf_nested_job> nj(ns);
  std::vector<tmf_minor_block> mb((size_t)c->nblocks);
  const double *e_dev = c->e_dev_ptr;
  if (e_dev == nullptr) return fail(TMF_ERR_VALUE, "tmf_chain_modes has not run");
  const int es = c->cplx ? 2 : 1;      // doubles per element of O / S / det / out (and per row of C_dev)
  int mb0 = 0;
  for (int u = 0; u < ns; ++u) {
    ChainSite &s = c->sites[u];
    const tmf_site_plan &h = s.plan.h;
    const int side = s.mode == 1 ? TMF_SIDE_R : TMF_SIDE_L;
    const ChainBond &bb = c->bonds[s.bra_bond], &kb = c->bonds[s.ket_bond];
    const ChainSide &bs = bb.side[side], &ks = kb.side[side];
    const int sb0 = h.s_bra - (h.ka_bra - h.k_always), sk0 = h.s_ket - (h.ka_ket - h.k_always);
    tmf_site_job &j = sj[u];
    std::memset(&j, 0, sizeof(j));
    j.Vb = V_dev + bs.v_off; j.Vk = V_dev + ks.v_off;
    j.ldb = std::max(bs.n, 1); j.ldk = std::max(ks.n, 1);
    j.bra_cols = s.cols_dev; j.ket_cols = s.cols_dev + tmf::PLAN_MAX_ORB;
    j.bra_sign = s.signs_dev; j.ket_sign = s.signs_dev + tmf::PLAN_MAX_ORB;
    j.O = O_dev + es * s.o_off; j.S = S_dev + es * s.s_off; j.det = det_dev + es * u;
    j.n_bra = h.n_bra; j.n_ket = h.n_ket; j.mode = h.mode; j.physical = h.physical;
    j.ka_bra = h.ka_bra; j.ka_ket = h.ka_ket; j.sb = sb0; j.sk = sk0;
    j.pad_[1] = 1;
    tmf_nested_job &q = nj[u];
    std::memset(&q, 0, sizeof(q));
    q.e_bra = e_dev + (size_t)bs.job * TMF_MAX_MODES;
    q.e_ket = e_dev + (size_t)ks.job * TMF_MAX_MODES;
    // (embedded complex matrix: row 2e holds conj(C[e, :]) = C[:, e]^T interleaved, entry [2e][2e] = C[e, e])
    q.c_edge = C_dev + (int64_t)es * s.site * ldc + es * s.site;
    q.a_col = (s.mode == 1) ? q.c_edge + es : C_dev + (int64_t)es * s.site * ldc;
    q.k_bra = bb.k; q.k_ket = kb.k; q.df = s.df;
    for (int b = 0; b < h.n_blocks; ++b) {
      const int *bl = &s.plan.blocks[6 * b];
      tmf_minor_block &k = mb[mb0 + b];
      std::memset(&k, 0, sizeof(k));
      k.S = j.S; k.det = j.det;
      k.bra_masks = s.bra_masks_dev + bl[0]; k.ket_masks = s.ket_masks_dev + bl[2];
      k.out = out_dev + es * s.block_off[b];
      k.s_bra = h.s_bra; k.s_ket = h.s_ket; k.n_bra = bl[1]; k.n_ket = bl[3]; k.minor = bl[4];
    }
    mb0 += h.n_blocks;
  }
  void *site_desc = ar.take<unsigned char>(tmf_site_desc_bytes(ns));
  void *minor_desc = ar.take<unsigned char>(tmf_minor_desc_bytes((int)mb.size()));
  if (!ar.ok()) return fail(TMF_ERR_VALUE, "plan workspace too small");
  int rc = c->cplx ? tmf_site_nested_c_batched(sj.data(), nj.data(), ns, site_desc, stream)
                   : tmf_site_nested_batched(sj.data(), nj.data(), ns, site_desc, stream);
  if (rc) return rc;
  if (c->peer_out && !mb.empty()) mb[0].pad_ |= 1;       // row-staged stores (NVLink-friendly)
  rc = c->cplx ? tmf_minors_blocks_c(mb.data(), (int)mb.size(), minor_desc, stream)
               : tmf_minors_blocks(mb.data(), (int)mb.size(), minor_desc, stream);
  if (rc) return rc;
  // bond tables -> pinned host staging, behind the kernels (read by the accessors / bulk exports)
  {
    std::lock_guard<std::mutex> lk(c->tab_mu);
    const size_t need = tmf::enumerate_tables_stage_bytes(c->enum_res);
    if (c->tab_cap < need) {
      g_pinned.release(c->tab_stage, c->tab_cap);
      c->tab_stage = g_pinned.acquire(need, c->tab_cap);
      if (!c->tab_stage) { c->tab_cap = 0; return fail(TMF_ERR_RUNTIME, "cannot allocate pinned staging memory"); }
    }
    rc = tmf::enumerate_fetch_tables(c->enum_res, c->tab_stage, stream);
    if (rc) return rc;
#if !defined(TMF_HOSTSIM)
    if (!c->tab_event) {
      cudaEvent_t ev;
      rc = tmf::check_cuda(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
      if (rc) return rc;
      c->tab_event = ev;
    }
    rc = tmf::check_cuda(cudaEventRecord((cudaEvent_t)c->tab_event, (cudaStream_t)stream), "cudaEventRecord");
    if (rc) return rc;
#endif
  }
  return TMF_OK;
}

int tmf_chain_tensors(tmf_chain *c, const double *C_dev, int ldc, double *V_dev, void *plan_dev,
                      int64_t plan_bytes, double *O_dev, double *S_dev, double *det_dev, double *out_dev,
                      void *stream) {
  if (!c->enumerated) return fail(TMF_ERR_VALUE, "tmf_chain_enumerate has not run");
  if (plan_bytes < c->plan_bytes) return fail(TMF_ERR_VALUE, "plan workspace too small");
  tmf::Arena ar(plan_dev, plan_bytes);
  StageTimer tm;
  int rc = centre_pairing(c, C_dev, ldc, V_dev, ar, stream);
  if (rc) return rc;
  tm.lap("tensors: centre pairing");
  if (c->device_plan) return chain_tensors_device_plan(c, C_dev, ldc, V_dev, ar, O_dev, S_dev, det_dev, out_dev, stream);
  const int es = c->cplx ? 2 : 1;      // doubles per element of O / S / det / out
  // ---- one blob with every per-site index / sign / mask array ------------------------------
  // pass 1 (serial, cheap): offsets of every array; pass 2 (threads over sites): copy + descriptors
  const int ns = (int)c->sites.size();
  unsigned char *blob_dev = ar.take<unsigned char>(0);
  struct SiteOff { size_t bc, kc, bs, ks, bm, km; int mb0; };
  std::vector<SiteOff> so(ns);
  size_t bsz = 0;
  int nmb = 0;
  auto reserve = [&](size_t bytes) {
    size_t off = (bsz + 255) & ~size_t(255);
    bsz = off + bytes;
    return off;
  };
  for (int u = 0; u < ns; ++u) {
    const tmf_site_plan &h = c->sites[u].plan.h;
    const int sb0 = h.s_bra - (h.ka_bra - h.k_always), sk0 = h.s_ket - (h.ka_ket - h.k_always);
    const size_t rows = (size_t)(h.ka_bra + sb0), cols = (size_t)(h.ka_ket + sk0);
    so[u].bc = reserve(4 * rows); so[u].kc = reserve(4 * cols);
    so[u].bs = reserve(8 * rows); so[u].ks = reserve(8 * cols);
    so[u].bm = reserve(8 * (size_t)h.n_rows); so[u].km = reserve(8 * (size_t)h.chi_ket);
    so[u].mb0 = nmb;
    nmb += h.n_blocks;
  }
  if (c->blob_cap < bsz) {
    g_pinned.release(c->blob, c->blob_cap);
    c->blob = g_pinned.acquire(bsz, c->blob_cap);
    if (!c->blob) { c->blob_cap = 0; return fail(TMF_ERR_RUNTIME, "cannot allocate pinned staging memory"); }
  }
  unsigned char *blob = c->blob;
  std::vector<tmf_site_job> sj(ns);
  std::vector<tmf_minor_block> mb(nmb);
  parallel_for(ns, c->n_threads, [&](int u) {
    ChainSite &s = c->sites[u];
    const tmf_site_plan &h = s.plan.h;
    const int side = s.mode == 1 ? TMF_SIDE_R : TMF_SIDE_L;
    const ChainSide &bs = c->bonds[s.bra_bond].side[side], &ks = c->bonds[s.ket_bond].side[side];
    const int sb0 = h.s_bra - (h.ka_bra - h.k_always), sk0 = h.s_ket - (h.ka_ket - h.k_always);
    const int rows = h.ka_bra + sb0, cols = h.ka_ket + sk0;
    auto put = [&](size_t off, const void *p, size_t bytes) {
      if (bytes) std::memcpy(blob + off, p, bytes);
      return blob_dev + off;
    };
    tmf_site_job &j = sj[u];
    std::memset(&j, 0, sizeof(j));
    j.Vb = V_dev + bs.v_off; j.Vk = V_dev + ks.v_off;
    j.ldb = std::max(bs.n, 1); j.ldk = std::max(ks.n, 1);
    j.bra_cols = reinterpret_cast<const int *>(put(so[u].bc, s.plan.bra_cols.data(), 4 * (size_t)rows));
    j.ket_cols = reinterpret_cast<const int *>(put(so[u].kc, s.plan.ket_cols.data(), 4 * (size_t)cols));
    j.bra_sign = reinterpret_cast<const double *>(put(so[u].bs, s.plan.bra_sign.data(), 8 * (size_t)rows));
    j.ket_sign = reinterpret_cast<const double *>(put(so[u].ks, s.plan.ket_sign.data(), 8 * (size_t)cols));
    j.O = O_dev + es * s.o_off; j.S = S_dev + es * s.s_off; j.det = det_dev + es * u;
    j.n_bra = h.n_bra; j.n_ket = h.n_ket; j.mode = h.mode; j.physical = h.physical;
    j.ka_bra = h.ka_bra; j.ka_ket = h.ka_ket; j.sb = sb0; j.sk = sk0;
    j.pad_[1] = 1;   // report a singular always block (incompatible neighbouring bonds) as NaN
    const uint64_t *bm = reinterpret_cast<const uint64_t *>(put(so[u].bm, s.plan.bra_masks.data(), 8 * (size_t)h.n_rows));
    const uint64_t *km = reinterpret_cast<const uint64_t *>(put(so[u].km, s.plan.ket_masks.data(), 8 * (size_t)h.chi_ket));
    for (int b = 0; b < h.n_blocks; ++b) {
      const int *bl = &s.plan.blocks[6 * b];
      tmf_minor_block &k = mb[so[u].mb0 + b];
      std::memset(&k, 0, sizeof(k));
      k.S = j.S; k.det = j.det;
      k.bra_masks = bm + bl[0]; k.ket_masks = km + bl[2];
      k.out = out_dev + es * s.block_off[b];
      k.s_bra = h.s_bra; k.s_ket = h.s_ket; k.n_bra = bl[1]; k.n_ket = bl[3]; k.minor = bl[4];
    }
  });
  ar.take<unsigned char>((int64_t)bsz);
  void *site_desc = ar.take<unsigned char>(tmf_site_desc_bytes(ns));
  void *minor_desc = ar.take<unsigned char>(tmf_minor_desc_bytes((int)mb.size()));
  if (!ar.ok()) return fail(TMF_ERR_VALUE, "plan workspace too small");
  tm.lap("tensors: build blob");
  rc = tmf::copy_h2d(blob_dev, blob, bsz, stream);
  if (rc) return rc;
  tm.lap("tensors: upload blob");
  if (c->nested) {
    std::vector<tmf_nested_job> nj(ns);
    const double *e_dev = c->e_dev_ptr;
    if (e_dev == nullptr) return fail(TMF_ERR_VALUE, "tmf_chain_modes has not run");
    for (int u = 0; u < ns; ++u) {
      const ChainSite &s = c->sites[u];
      const int side = s.mode == 1 ? TMF_SIDE_R : TMF_SIDE_L;
      const ChainBond &bb = c->bonds[s.bra_bond], &kb = c->bonds[s.ket_bond];
      tmf_nested_job &q = nj[u];
      std::memset(&q, 0, sizeof(q));
      q.e_bra = e_dev + (size_t)bb.side[side].job * TMF_MAX_MODES;
      q.e_ket = e_dev + (size_t)kb.side[side].job * TMF_MAX_MODES;
      q.c_edge = C_dev + (int64_t)es * s.site * ldc + es * s.site;
      q.a_col = (s.mode == 1) ? q.c_edge + es : C_dev + (int64_t)es * s.site * ldc;
      q.k_bra = bb.k; q.k_ket = kb.k; q.df = s.df;
    }
    rc = c->cplx ? tmf_site_nested_c_batched(sj.data(), nj.data(), ns, site_desc, stream)
                 : tmf_site_nested_batched(sj.data(), nj.data(), ns, site_desc, stream);
  } else {
    rc = tmf_site_overlap_schur_batched(sj.data(), ns, site_desc, stream);
  }
  if (rc) return rc;
  tm.lap("tensors: enqueue site kernels");
  if (c->peer_out && !mb.empty()) mb[0].pad_ |= 1;
  rc = c->cplx ? tmf_minors_blocks_c(mb.data(), (int)mb.size(), minor_desc, stream)
               : tmf_minors_blocks(mb.data(), (int)mb.size(), minor_desc, stream);
  tm.lap("tensors: enqueue minors");
  return rc;
}

// ---- result accessors (host pointers stay valid until tmf_chain_destroy) ----------------------
int tmf_chain_bond(tmf_chain *c, int bond, int *q, const double **lam, const int **charge,
                   const uint64_t **masks, const int **sec_q, const int **sec_start,
                   const double **e) {
  if (bond < 0 || bond > c->L || !c->bonds[bond].used || !c->enumerated)
    return fail(TMF_ERR_VALUE, "bond not available on this shard");
  if (int rc = chain_ensure_tables(c)) return rc;
  const ChainBond &B = c->bonds[bond];
  q[0] = (int)B.bv.masks.size(); q[1] = B.k; q[2] = B.filled_left; q[3] = (int)B.bv.sec_q.size();
  q[4] = B.side[0].job; q[5] = B.side[1].job; q[6] = B.side[0].f; q[7] = B.side[1].f;
  *lam = B.bv.lam.data(); *charge = B.bv.charge.data(); *masks = B.bv.masks.data();
  *sec_q = B.bv.sec_q.data(); *sec_start = B.bv.sec_start.data();
  const int job = (B.side[0].job >= 0) ? B.side[0].job : B.side[1].job;
  *e = c->e_host.data() + (size_t)job * TMF_MAX_MODES;
  return TMF_OK;
}

int tmf_chain_site(tmf_chain *c, int site, tmf_site_plan *plan, const int **blocks,
                   const int64_t **block_off, const int **row_p, const int **row_alpha,
                   int64_t *offs) {
  if (site < c->site_lo || site >= c->site_hi || !c->enumerated)
    return fail(TMF_ERR_VALUE, "site not available on this shard");
  if (int rc = chain_ensure_tables(c)) return rc;
  chain_site_rows(c, c->sites[site - c->site_lo]);
  const ChainSite &s = c->sites[site - c->site_lo];
  *plan = s.plan.h;
  *blocks = s.plan.blocks.data(); *block_off = s.block_off.data();
  *row_p = s.plan.row_p.data(); *row_alpha = s.plan.row_alpha.data();
  offs[0] = s.o_off; offs[1] = s.s_off; offs[2] = site - c->site_lo;
  return TMF_OK;
}

// ---- bulk export: every bond / site table of the shard in one call each (the per-object accessors
// above cost one FFI round trip per bond / site, which dominated the end-to-end time at L = 1024) ----
// q = {first bond, number of bonds, sum of chi, sum of sector counts}
int tmf_chain_bonds_sizes(tmf_chain *c, int64_t *q) {
  if (!c->enumerated) return fail(TMF_ERR_VALUE, "tmf_chain_enumerate has not run");
  if (int rc = chain_ensure_tables(c)) return rc;
  int first = -1, last = -1;
  int64_t chi = 0, nsec = 0;
  for (int b = 0; b <= c->L; ++b)
    if (c->bonds[b].used) {
      if (first < 0) first = b;
      last = b;
      chi += (int64_t)c->bonds[b].bv.masks.size();
      nsec += (int64_t)c->bonds[b].bv.sec_q.size();
    }
  // bonds of a shard are contiguous except for the centre bond pulled in by the pairing
  q[0] = first; q[1] = (first < 0) ? 0 : last - first + 1; q[2] = chi; q[3] = nsec;
  return TMF_OK;
}

// Arrays are indexed by (bond - first); unused bonds in the range get chi = 0.
//   chi_off[nb + 1], head[4 nb] = {k, filled_left, fL, fR}, lam / charge / masks [sum chi],
//   sec_off[nb + 1], sec_q[sum nsec], sec_start[sum nsec + nb] (nsec + 1 entries per bond), e[64 nb]
int tmf_chain_bonds_export(tmf_chain *c, int64_t *chi_off, int *head, double *lam, int *charge,
                           uint64_t *masks, int64_t *sec_off, int *sec_q, int *sec_start, double *e) {
  int64_t q[4];
  int rc = tmf_chain_bonds_sizes(c, q);
  if (rc) return rc;
  const int first = (int)q[0], nb = (int)q[1];
  int64_t co = 0, so = 0;
  for (int i = 0; i < nb; ++i) {
    const ChainBond &B = c->bonds[first + i];
    chi_off[i] = co;
    sec_off[i] = so;
    int *h = head + 4 * i;
    h[0] = h[1] = h[2] = h[3] = 0;
    double *ei = e + (size_t)i * TMF_MAX_MODES;
    std::memset(ei, 0, sizeof(double) * TMF_MAX_MODES);
    if (!B.used) { sec_start[so + i] = 0; continue; }
    const size_t chi = B.bv.masks.size(), ns = B.bv.sec_q.size();
    h[0] = B.k; h[1] = B.filled_left; h[2] = B.side[0].f; h[3] = B.side[1].f;
    if (chi) {
      std::memcpy(lam + co, B.bv.lam.data(), sizeof(double) * chi);
      std::memcpy(charge + co, B.bv.charge.data(), sizeof(int) * chi);
      std::memcpy(masks + co, B.bv.masks.data(), sizeof(uint64_t) * chi);
    }
    if (ns) std::memcpy(sec_q + so, B.bv.sec_q.data(), sizeof(int) * ns);
    if (B.bv.sec_start.size() == ns + 1) std::memcpy(sec_start + so + i, B.bv.sec_start.data(), sizeof(int) * (ns + 1));
    else sec_start[so + i] = 0;
    const int job = (B.side[0].job >= 0) ? B.side[0].job : B.side[1].job;
    if (job >= 0 && B.k > 0) std::memcpy(ei, c->e_host.data() + (size_t)job * TMF_MAX_MODES, sizeof(double) * std::min(B.k, (int)TMF_MAX_MODES));
    co += (int64_t)chi;
    so += (int64_t)ns;
  }
  chi_off[nb] = co;
  sec_off[nb] = so;
  return TMF_OK;
}

// q = {number of sites, sum of n_blocks, sum of n_rows}
int tmf_chain_sites_sizes(tmf_chain *c, int64_t *q) {
  if (!c->enumerated) return fail(TMF_ERR_VALUE, "tmf_chain_enumerate has not run");
  int64_t nbk = 0, nr = 0;
  for (const ChainSite &s : c->sites) { nbk += s.plan.h.n_blocks; nr += s.plan.h.n_rows; }
  q[0] = (int64_t)c->sites.size(); q[1] = nbk; q[2] = nr;
  return TMF_OK;
}

//   plans[ns], blk_off[ns + 1], blocks[6 sum n_blocks], block_off[sum n_blocks] (element offsets into the
//   shard's tensor buffer), row_off[ns + 1], row_p / row_alpha [sum n_rows]
int tmf_chain_sites_export(tmf_chain *c, tmf_site_plan *plans, int64_t *blk_off, int *blocks,
                           int64_t *block_off, int64_t *row_off, int *row_p, int *row_alpha) {
  if (!c->enumerated) return fail(TMF_ERR_VALUE, "tmf_chain_enumerate has not run");
  if (row_p != nullptr && row_alpha != nullptr)
    if (int rc = chain_ensure_tables(c)) return rc;
  const int ns = (int)c->sites.size();
  int64_t bo = 0, ro = 0;
  for (int u = 0; u < ns; ++u) { blk_off[u] = bo; row_off[u] = ro; bo += c->sites[u].plan.h.n_blocks; ro += c->sites[u].plan.h.n_rows; }
  blk_off[ns] = bo;
  row_off[ns] = ro;
  parallel_for(ns, c->n_threads > 0 ? c->n_threads : 4, [&](int u) {
    if (row_p != nullptr && row_alpha != nullptr) chain_site_rows(c, c->sites[u]);
    const ChainSite &s = c->sites[u];
    const tmf_site_plan &h = s.plan.h;
    plans[u] = h;
    if (h.n_blocks) {
      std::memcpy(blocks + 6 * blk_off[u], s.plan.blocks.data(), sizeof(int) * 6 * (size_t)h.n_blocks);
      std::memcpy(block_off + blk_off[u], s.block_off.data(), sizeof(int64_t) * (size_t)h.n_blocks);
    }
    if (h.n_rows && row_p != nullptr && row_alpha != nullptr) {
      std::memcpy(row_p + row_off[u], s.plan.row_p.data(), sizeof(int) * (size_t)h.n_rows);
      std::memcpy(row_alpha + row_off[u], s.plan.row_alpha.data(), sizeof(int) * (size_t)h.n_rows);
    }
  });
  return TMF_OK;
}

int64_t tmf_chain_job_voff(tmf_chain *c, int job) { return c->v_off[job]; }

// Algorithmic FP64 flop counts of the *reference's* algorithm for this shard:
// f[0] eigh (10/3 n^3 per job), f[1] overlap GEMM 2 (n+1) c_b c_k, f[2] Schur (8/3) k^3,
// f[3] minors sum nsb nsk (2/3) q^3, f[4] number of minors.
int tmf_chain_flops(tmf_chain *c, double *f) {
  if (!c->enumerated) return fail(TMF_ERR_VALUE, "tmf_chain_enumerate has not run");
  for (int i = 0; i < 5; ++i) f[i] = 0.0;
  for (size_t j = 0; j < c->job_x.size(); ++j) {
    double n = (c->job_side[j] == TMF_SIDE_L) ? c->job_x[j] : c->L - c->job_x[j];
    f[0] += 10.0 / 3.0 * n * n * n;
  }
  for (const ChainSite &s : c->sites) {
    const tmf_site_plan &h = s.plan.h;
    // (the reference's O and always block contain the filled orbitals the nested form eliminates in closed form)
    const double cb = h.ka_bra + (h.s_bra - (h.ka_bra - h.k_always)) + s.f_common;
    const double ck = h.ka_ket + (h.s_ket - (h.ka_ket - h.k_always)) + s.f_common;
    const double ka = (double)h.k_always + s.f_common;
    f[1] += 2.0 * (h.n_bra + 1.0) * cb * ck;
    f[2] += 8.0 / 3.0 * ka * ka * ka;
    for (int b = 0; b < h.n_blocks; ++b) {
      const int *bl = &s.plan.blocks[6 * b];
      f[3] += (double)bl[1] * bl[3] * (2.0 / 3.0) * bl[4] * bl[4] * bl[4];
      f[4] += (double)bl[1] * bl[3];
    }
  }
  return TMF_OK;
}

}  // extern "C"
