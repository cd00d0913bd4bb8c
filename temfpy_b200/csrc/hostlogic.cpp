// Host-side integer / ordering logic.  See hostlogic.hpp.
// Compiled with -ffp-contract=off so that the running sums are formed exactly like the
// reference's Python floats (one IEEE add per step).
#include "hostlogic.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <queue>
#include <stdexcept>
#include <thread>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <exception>
#include <memory>
#include <mutex>

namespace tmf {

// Last error message.  Every thread keeps its own (the pipeline drives the library from several threads at
// once); the process-wide copy serves callers that read the message from another thread than the failing one
// and is guarded by a mutex (concurrent failures of several chunks -- e.g. the sketch-width retry -- would
// otherwise race on the string).  The pointer handed out always refers to a thread-local buffer.
static thread_local std::string g_error;
static thread_local std::string g_error_out;
static std::mutex g_error_mu;
static std::string g_error_shared;
void set_error(const std::string &msg) {
  g_error = msg;
  std::lock_guard<std::mutex> lk(g_error_mu);
  g_error_shared = msg;
}
const char *last_error_cstr() {
  if (!g_error.empty()) {
    g_error_out = g_error;
  } else {
    std::lock_guard<std::mutex> lk(g_error_mu);
    g_error_out = g_error_shared;
  }
  return g_error_out.c_str();
}

// -------------------------------------------------------------------------------------------
// persistent host thread pool
// -------------------------------------------------------------------------------------------
namespace {
struct PoolJob {
  std::atomic<int> next{0}, finished{0};
  int n = 0;
  const std::function<void(int)> *f = nullptr;
  std::mutex mu;
  std::condition_variable cv;
  std::exception_ptr err;
  void work() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) break;
      try {
        (*f)(i);
      } catch (...) {
        std::lock_guard<std::mutex> lk(mu);
        if (!err) err = std::current_exception();
      }
      if (finished.fetch_add(1) + 1 == n) {
        std::lock_guard<std::mutex> lk(mu);
        cv.notify_all();
      }
    }
  }
};
struct Pool {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::shared_ptr<PoolJob>> q;
  std::vector<std::thread> workers;
  bool stop = false;
  int size = 0;
  Pool() {
    size = (int)std::max(1u, std::thread::hardware_concurrency());
    for (int t = 0; t < size; ++t)
      workers.emplace_back([this] {
        for (;;) {
          std::shared_ptr<PoolJob> job;
          {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [this] { return stop || !q.empty(); });
            if (stop && q.empty()) return;
            job = q.front();
            q.pop_front();
          }
          job->work();
        }
      });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    for (auto &w : workers) w.join();
  }
};
Pool &pool() {
  static Pool *p = new Pool();   // intentionally leaked: worker threads must outlive static destruction order
  return *p;
}
}  // namespace

void pool_for(int n, int max_threads, const std::function<void(int)> &f) {
  if (n <= 0) return;
  if (max_threads <= 0) max_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  if (n == 1 || max_threads == 1) {
    for (int i = 0; i < n; ++i) f(i);
    return;
  }
  Pool &p = pool();
  auto job = std::make_shared<PoolJob>();
  job->n = n;
  job->f = &f;
  const int helpers = std::min(std::min(max_threads, n), p.size) - 1;
  {
    std::lock_guard<std::mutex> lk(p.mu);
    for (int h = 0; h < helpers; ++h) p.q.push_back(job);
  }
  if (helpers > 0) p.cv.notify_all();
  job->work();
  {
    std::unique_lock<std::mutex> lk(job->mu);
    job->cv.wait(lk, [&] { return job->finished.load() >= n; });
  }
  if (job->err) std::rethrow_exception(job->err);
}

bool TruncPar::is_sector(int q) const {
  if (!filter) return true;
  return std::find(sectors.begin(), sectors.end(), q) != sectors.end();
}

namespace {

struct HeapItem {
  double sum;
  int64_t seq;
  int i;
  uint64_t set;
};
struct HeapCmp {  // min-heap on (sum, seq): same order as Python tuples (schmidt_utils.py:293)
  bool operator()(const HeapItem &x, const HeapItem &y) const {
    if (x.sum != y.sum) return x.sum > y.sum;
    return x.seq > y.seq;
  }
};

inline int charge_of(uint64_t set, int k, int filled_left, int filled_right) {
  int n = __builtin_popcountll(set);
  if (filled_left >= 0) return filled_left + n;          // schmidt_utils.py:266
  if (filled_right >= 0) return filled_right + k - n;    // :264
  return n;                                              // :262
}

// StoppingCondition.__call__ (schmidt_utils.py:99-138)
inline bool more_needed(const std::vector<double> &s, const TruncPar &tp, double max_logval) {
  if (tp.chi_max >= 0 && (int)s.size() > tp.chi_max) return false;
  if (s.back() - s.front() > max_logval) return false;
  return true;
}

// StoppingCondition.truncate (schmidt_utils.py:140-185)
int truncate(const std::vector<double> &lv, const TruncPar &tp) {
  const int n = (int)lv.size();
  const double lim = -std::log(tp.svd_min);
  int cut = -1;
  for (int i = 0; i < n; ++i) {
    bool ok = true;
    if (tp.chi_max >= 0 && i >= tp.chi_max) ok = false;
    if (!(lv[i] - lv[0] < lim)) ok = false;
    if (i < n - 1 && !((lv[i + 1] - lv[i]) > tp.degeneracy_tol)) ok = false;
    if (ok) cut = i;
  }
  if (cut < 0) throw std::runtime_error("truncate: no admissible cut");  // reference: IndexError
  return cut + 1;
}

}  // namespace

namespace {
// Monotone priority queue for the best-first enumeration.  Popped sums never decrease (a child is never
// smaller than its parent, up to one rounding) and only the window [base, base + max_logval] matters, so
// the items are binned by sum into a few thousand buckets: a pop takes the (sum, seq)-minimum of the lowest
// non-empty bucket (a scan of one or two items), a push is an index computation.  The order of the pops --
// and with it every result -- is exactly that of the reference's binary heap; it is just ~4x cheaper than
// sifting through a 1000-entry heap.  Items beyond the window go to a small binary heap.
struct BucketQueue {
  static constexpr int NB = 8192;
  std::vector<int32_t> head, next;
  std::vector<uint64_t> bits;
  std::vector<HeapItem> pool, over;
  double base = 0.0, scale = 0.0;
  int cur = NB;
  static bool less(const HeapItem &x, const HeapItem &y) {
    return (x.sum < y.sum) | ((x.sum == y.sum) & (x.seq < y.seq));
  }
  void reset(double base_, double window) {
    head.assign(NB, -1);
    bits.assign(NB / 64, 0);
    pool.clear();
    next.clear();
    over.clear();
    base = base_;
    scale = NB / (window * 1.0001);
    cur = NB;
  }
  void push(const HeapItem &v) {
    const double t = (v.sum - base) * scale;
    if (!(t < (double)NB)) {          // beyond the window (or NaN): overflow heap
      size_t i = over.size();
      over.push_back(v);
      while (i > 0) {
        size_t p = (i - 1) / 2;
        if (!less(v, over[p])) break;
        over[i] = over[p];
        i = p;
      }
      over[i] = v;
      return;
    }
    const int b = t > 0.0 ? (int)t : 0;
    const int32_t id = (int32_t)pool.size();
    pool.push_back(v);
    next.push_back(head[b]);
    head[b] = id;
    bits[b >> 6] |= (1ull << (b & 63));
    if (b < cur) cur = b;
  }
  bool empty() {
    if (cur < NB && head[cur] >= 0) return false;
    int w = cur >> 6;
    if (cur < NB) {
      uint64_t m = bits[w] & (~0ull << (cur & 63));
      for (;;) {
        if (m) { cur = (w << 6) + __builtin_ctzll(m); return false; }
        if (++w >= NB / 64) break;
        m = bits[w];
      }
      cur = NB;
    }
    return over.empty();
  }
  HeapItem pop() {   // precondition: !empty()
    if (cur < NB) {
      int32_t best = head[cur], bprev = -1, prev = head[cur];
      for (int32_t it = next[prev]; it >= 0; prev = it, it = next[it])
        if (less(pool[it], pool[best])) { best = it; bprev = prev; }
      if (bprev < 0) head[cur] = next[best]; else next[bprev] = next[best];
      if (head[cur] < 0) bits[cur >> 6] &= ~(1ull << (cur & 63));
      return pool[best];
    }
    const HeapItem top = over[0];
    const HeapItem v = over.back();
    over.pop_back();
    const size_t nh = over.size();
    if (nh) {
      size_t i = 0;
      for (;;) {
        size_t c = 2 * i + 1;
        if (c >= nh) break;
        if (c + 1 < nh && less(over[c + 1], over[c])) ++c;
        if (!less(over[c], v)) break;
        over[i] = over[c];
        i = c;
      }
      over[i] = v;
    }
    return top;
  }
};
}  // namespace

void lowest_sums(const double *a, int k, double base, const TruncPar &tp, int filled_left,
                 int filled_right, std::vector<double> &sums, std::vector<uint64_t> &sets,
                 int *n_checked) {
  sums.clear();
  sets.clear();
  if (k > TMF_MAX_MODES) throw std::invalid_argument("more than 64 entangled modes on one bond");
  if (k == 0) {  // schmidt_utils.py:268-271
    if (tp.is_sector(charge_of(0, 0, filled_left, filled_right))) {
      sums.push_back(0.0);
      sets.push_back(0);
    }
    if (n_checked) *n_checked = 1;
    return;
  }
  const double max_logval = -std::log(tp.svd_min) + tp.degeneracy_tol;  // :96
  uint64_t neg = 0;
  for (int i = 0; i < k; ++i)
    if (a[i] < 0) neg |= (1ull << i);
  if (tp.is_sector(charge_of(neg, k, filled_left, filled_right))) {  // :277-279
    sums.push_back(base);
    sets.push_back(neg);
  }
  std::vector<double> mag(k);
  std::vector<int> order(k);
  for (int i = 0; i < k; ++i) {
    mag[i] = std::fabs(a[i]);
    order[i] = i;
  }
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return mag[x] < mag[y]; });
  // binary min-heap on (sum, seq) in a flat vector; "replace top" fuses the pop with the first push
  std::vector<HeapItem> heap;
  heap.reserve(tp.chi_max > 0 ? 2 * (size_t)tp.chi_max + 8 : 4096);
  // (sum, seq) order without short-circuit branches: the comparison outcomes along a sift are unpredictable
  auto less = [](const HeapItem &x, const HeapItem &y) {
    return (x.sum < y.sum) | ((x.sum == y.sum) & (x.seq < y.seq));
  };
  // Replacing the root: the new item is a child of the popped one (a larger sum), so it belongs near the
  // bottom -- walk the hole down along the smaller children without comparing against it, then sift it up.
  auto sift_down = [&](size_t i) {
    const size_t nh = heap.size();
    const HeapItem v = heap[i];
    for (;;) {
      size_t c = 2 * i + 1;
      if (c >= nh) break;
      if (c + 1 < nh) c += (size_t)less(heap[c + 1], heap[c]);
      if (!less(heap[c], v)) break;
      heap[i] = heap[c];
      i = c;
    }
    heap[i] = v;
  };
  auto push = [&](const HeapItem &v) {
    size_t i = heap.size();
    heap.push_back(v);
    while (i > 0) {
      size_t p = (i - 1) / 2;
      if (!less(v, heap[p])) break;
      heap[i] = heap[p];
      i = p;
    }
    heap[i] = v;
  };
  int64_t seq = 0;
  int checked = 1;
  if (max_logval > 0.0 && max_logval < 64.0) {
    // fast path: bucket queue over the window of admissible sums
    static thread_local BucketQueue q;
    q.reset(base, max_logval);
    q.push({base + mag[order[0]], seq, 0, neg ^ (1ull << order[0])});  // :291-293
    while (!q.empty() && (sums.empty() || more_needed(sums, tp, max_logval))) {  // :297
      ++checked;
      const HeapItem it = q.pop();
      if (tp.is_sector(charge_of(it.set, k, filled_left, filled_right))) {
        sums.push_back(it.sum);
        sets.push_back(it.set);
      }
      if (it.i < k - 1) {  // :304-315
        const uint64_t c1 = it.set ^ (1ull << order[it.i + 1]);
        double s = it.sum + mag[order[it.i + 1]];
        q.push({s, ++seq, it.i + 1, c1});
        const uint64_t c2 = c1 ^ (1ull << order[it.i]);
        s = s - mag[order[it.i]];
        q.push({s, ++seq, it.i + 1, c2});
      }
    }
  } else {
  push({base + mag[order[0]], seq, 0, neg ^ (1ull << order[0])});  // :291-293
  while (!heap.empty() && (sums.empty() || more_needed(sums, tp, max_logval))) {  // :297
    ++checked;
    const HeapItem it = heap[0];
    if (tp.is_sector(charge_of(it.set, k, filled_left, filled_right))) {
      sums.push_back(it.sum);
      sets.push_back(it.set);
    }
    if (it.i < k - 1) {  // :304-315
      uint64_t c1 = it.set ^ (1ull << order[it.i + 1]);
      double s = it.sum + mag[order[it.i + 1]];
      heap[0] = {s, ++seq, it.i + 1, c1};   // heappop + first heappush
      sift_down(0);
      uint64_t c2 = c1 ^ (1ull << order[it.i]);
      s = s - mag[order[it.i]];
      push({s, ++seq, it.i + 1, c2});
    } else {
      heap[0] = heap.back();
      heap.pop_back();
      if (!heap.empty()) sift_down(0);
    }
  }
  }
  if (n_checked) *n_checked = checked;
  if (sums.empty()) return;
  int cut = truncate(sums, tp);  // :321
  sums.resize(cut);
  sets.resize(cut);
}

// numpy's pairwise summation for n < 128 (used by np.sum in schmidt_utils.py:274)
static double numpy_sum(const double *v, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r += v[i];
    return r;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = v[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] += v[i + j];
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += v[i];
  return res;
}

// Mode weights a_i = log((1-e_i)/e_i)/2 amplify the rounding error of the eigenvalue by 1/(2 e (1-e)):
// for weak modes (e or 1-e ~ 1e-8) two eigenvalues that are equal by symmetry (spin partners of a
// spinful chain, particle-hole partners) come out with |a| differing by ~1e-8, far above
// degeneracy_tol, and the chi_max cut may then split their Schmidt multiplet -- which side survives
// is rounding noise (in the reference it is LAPACK's noise) and can leave neighbouring
// bonds with incompatible vector sets.  Weights that are indistinguishable within the eigenvalue
// accuracy (4e-15 absolute, the calibrated noise of the mode solver) are therefore set to their
// common mean before the enumeration, so that `truncate` sees the multiplets as the exact
// degeneracies they are.  Well-conditioned weights are untouched (tolerance 1.6e-14 at e = 1/2).
void snap_degenerate(double *a, const double *e, int k) {
  std::vector<int> ord(k);
  for (int i = 0; i < k; ++i) ord[i] = i;
  std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return std::fabs(a[x]) < std::fabs(a[y]); });
  auto tol = [&](int i) {
    const double w = e[i] * (1.0 - e[i]);
    return (w > 0.0) ? 4e-15 / (2.0 * w) : 0.0;
  };
  int s0 = 0;
  while (s0 < k) {
    int s1 = s0 + 1;
    while (s1 < k && std::fabs(a[ord[s1]]) - std::fabs(a[ord[s1 - 1]]) <=
                         std::max(tol(ord[s1]), tol(ord[s1 - 1])))
      ++s1;
    if (s1 - s0 > 1) {
      double mean = 0.0;
      for (int t = s0; t < s1; ++t) mean += std::fabs(a[ord[t]]);
      mean /= (s1 - s0);
      for (int t = s0; t < s1; ++t) a[ord[t]] = (a[ord[t]] < 0) ? -mean : mean;
    }
    s0 = s1;
  }
}

void bond_vectors(const double *e, int k, int filled_left, const TruncPar &tp, BondVectors &out) {
  out.k = k;
  out.filled_left = filled_left;
  std::vector<double> a(k), negs;
  for (int i = 0; i < k; ++i) {
    a[i] = std::log((1.0 - e[i]) / e[i]) / 2;  // slater.py:428, :663
    if (a[i] < 0) negs.push_back(a[i]);
  }
  if (tp.snap) snap_degenerate(a.data(), e, k);
  negs.clear();
  for (int i = 0; i < k; ++i)
    if (a[i] < 0) negs.push_back(a[i]);
  double base = numpy_sum(negs.data(), (int)negs.size());
  std::vector<double> sums;
  std::vector<uint64_t> sets;
  lowest_sums(a.data(), k, base, tp, filled_left, -1, sums, sets, nullptr);
  const int chi = (int)sets.size();
  if (chi == 0)
    throw std::invalid_argument("No Schmidt vectors left after filtering by `trunc_par.sectors`!");
  std::vector<int> idx(chi), nl(chi);
  for (int i = 0; i < chi; ++i) {
    idx[i] = i;
    nl[i] = filled_left + __builtin_popcountll(sets[i]);  // slater.py:673
  }
  {  // stable sort by n_L (slater.py:676): counting sort, the charges span at most k + 1 values
    std::vector<int> start((size_t)k + 2, 0);
    for (int i = 0; i < chi; ++i) ++start[nl[i] - filled_left + 1];
    for (int q = 1; q <= k + 1; ++q) start[q] += start[q - 1];
    for (int i = 0; i < chi; ++i) idx[start[nl[i] - filled_left]++] = i;
  }
  out.masks.resize(chi);
  out.lam.resize(chi);
  out.charge.resize(chi);
  out.sec_q.clear();
  out.sec_start.clear();
  // slater.py:489: lambda^2 = prod_i (e_i if occupied else 1 - e_i), multiplied in mode order.  Table lookup
  // instead of a branch per mode, four vectors at a time to overlap the multiply latencies.
  std::vector<double> tab(2 * (size_t)std::max(k, 1));
  for (int i = 0; i < k; ++i) { tab[2 * i] = 1.0 - e[i]; tab[2 * i + 1] = e[i]; }
  for (int r = 0; r < chi; ++r) {
    out.masks[r] = sets[idx[r]];
    out.charge[r] = nl[idx[r]];
  }
  int r4 = 0;
  for (; r4 + 4 <= chi; r4 += 4) {
    const uint64_t m0 = out.masks[r4], m1 = out.masks[r4 + 1], m2 = out.masks[r4 + 2], m3 = out.masks[r4 + 3];
    double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
    for (int i = 0; i < k; ++i) {
      p0 *= tab[2 * i + ((m0 >> i) & 1)];
      p1 *= tab[2 * i + ((m1 >> i) & 1)];
      p2 *= tab[2 * i + ((m2 >> i) & 1)];
      p3 *= tab[2 * i + ((m3 >> i) & 1)];
    }
    out.lam[r4] = std::sqrt(p0); out.lam[r4 + 1] = std::sqrt(p1);
    out.lam[r4 + 2] = std::sqrt(p2); out.lam[r4 + 3] = std::sqrt(p3);
  }
  for (; r4 < chi; ++r4) {
    const uint64_t m = out.masks[r4];
    double p = 1.0;
    for (int i = 0; i < k; ++i) p *= tab[2 * i + ((m >> i) & 1)];
    out.lam[r4] = std::sqrt(p);
  }
  for (int r = 0; r < chi; ++r)
    if (r == 0 || out.charge[r] != out.charge[r - 1]) {
      out.sec_q.push_back(out.charge[r]);
      out.sec_start.push_back(r);
    }
  out.sec_start.push_back(chi);
}

// -------------------------------------------------------------------------------------------
// per-site planning
// -------------------------------------------------------------------------------------------
namespace {
enum Kind { ENT = 0, FILLED = 1, PHYS = 2 };
struct Orb {
  int kind, idx;  // ENT: mode index; FILLED: index within the filled basis
  int cls;        // 0 never, 1 sometimes, 2 always
  int col;        // stored-V column (-1 physical)
};

// orbitals of one bond side in the reference's column order (slater.py:355-368, :1030-1051)
std::vector<Orb> side_orbitals(int mode, int k, int f, int chi, const uint64_t *masks,
                               bool with_phys) {
  uint64_t all = ~0ull, any = 0;
  for (int i = 0; i < chi; ++i) {
    all &= masks[i];
    any |= masks[i];
  }
  auto ent_cls = [&](int i) {
    bool a = (all >> i) & 1, y = (any >> i) & 1;  // occupation on the LEFT
    if (mode == 0) return a ? 2 : (y ? 1 : 0);    // left vectors: occupied iff bit set
    return !y ? 2 : (a ? 0 : 1);                  // right vectors: occupied iff bit clear
  };
  std::vector<Orb> o;
  if (mode == 0) {
    for (int t = 0; t < f; ++t) o.push_back({FILLED, t, 2, k + t});
    for (int i = 0; i < k; ++i) o.push_back({ENT, i, ent_cls(i), i});
    if (with_phys) o.push_back({PHYS, 0, 1, -1});
  } else {
    if (with_phys) o.push_back({PHYS, 0, 1, -1});
    for (int i = k - 1; i >= 0; --i) o.push_back({ENT, i, ent_cls(i), i});
    for (int t = 0; t < f; ++t) o.push_back({FILLED, t, 2, k + t});
  }
  return o;
}

struct Selected {
  std::vector<Orb> always, sometimes;
  std::vector<double> sometimes_sign;
};
// slater.py:792-821
Selected select(const std::vector<Orb> &orbs, int mode) {
  Selected s;
  int n_before = 0;
  std::vector<int> before;
  for (const Orb &o : orbs) {
    if (o.cls == 2) {
      s.always.push_back(o);
      ++n_before;
    } else if (o.cls == 1) {
      s.sometimes.push_back(o);
      before.push_back(n_before);
    }
  }
  const int k = (int)s.always.size();
  for (int b : before) {
    int expo = (mode == 0) ? (k - b) : b;
    s.sometimes_sign.push_back((expo & 1) ? -1.0 : 1.0);
  }
  return s;
}
}  // namespace

void site_plan(int mode, int n_bra, int n_ket, int k_bra, int f_bra, int nferm_bra, int chi_bra,
               const uint64_t *masks_bra, const int *charge_bra, int k_ket, int f_ket,
               int nferm_ket, int chi_ket, const uint64_t *masks_ket, const int *charge_ket,
               SitePlan &out) {
  bool physical;
  if (n_bra == n_ket)
    physical = false;
  else if (n_bra + 1 == n_ket)
    physical = true;
  else
    throw std::invalid_argument("`Schmidt_bra` must match or be one bond shorter than `Schmidt_ket`");
  if (chi_bra <= 0 || chi_ket <= 0) throw std::invalid_argument("empty bond");

  std::vector<Orb> ob = side_orbitals(mode, k_bra, f_bra, chi_bra, masks_bra, physical);
  std::vector<Orb> ok = side_orbitals(mode, k_ket, f_ket, chi_ket, masks_ket, false);
  Selected sb = select(ob, mode), sk = select(ok, mode);
  const int kb = (int)sb.always.size(), kk = (int)sk.always.size();
  const int k = std::min(kb, kk);  // slater.py:1069

  // canonical arrangement: [always block (k) | sometimes part], sometimes part in reference order
  struct Entry {
    Orb o;
    double sign;
  };
  auto arrange = [&](const Selected &s, int kside, std::vector<Entry> &blk, std::vector<Entry> &rest) {
    if (mode == 0) {  // idx = (always, sometimes); block = first k; rest = always[k:] + sometimes
      for (int i = 0; i < k; ++i) blk.push_back({s.always[i], 1.0});
      for (int i = k; i < kside; ++i) rest.push_back({s.always[i], 1.0});
      for (size_t i = 0; i < s.sometimes.size(); ++i)
        rest.push_back({s.sometimes[i], s.sometimes_sign[i]});
    } else {  // idx = (sometimes, always); block = last k; rest = sometimes + always[:kside-k]
      for (size_t i = 0; i < s.sometimes.size(); ++i)
        rest.push_back({s.sometimes[i], s.sometimes_sign[i]});
      for (int i = 0; i < kside - k; ++i) rest.push_back({s.always[i], 1.0});
      for (int i = kside - k; i < kside; ++i) blk.push_back({s.always[i], 1.0});
    }
  };
  std::vector<Entry> bb, br, kb_, kr;
  arrange(sb, kb, bb, br);
  arrange(sk, kk, kb_, kr);
  if (br.size() > 64 || kr.size() > 64)
    throw std::invalid_argument("sometimes matrix larger than 64 rows/cols is not supported");

  // device arrangement of O: [all always orbitals | sometimes orbitals]; which always orbitals form
  // the square block is decided by pivoting on the device (any choice gives the same tensor up to a
  // global sign), the surplus ones join the sometimes part as all-occupied rows / columns.
  out.bra_cols.clear(); out.bra_sign.clear(); out.ket_cols.clear(); out.ket_sign.clear();
  for (auto &o : sb.always) { out.bra_cols.push_back(o.col); out.bra_sign.push_back(1.0); }
  for (size_t i = 0; i < sb.sometimes.size(); ++i) {
    out.bra_cols.push_back(sb.sometimes[i].col); out.bra_sign.push_back(sb.sometimes_sign[i]);
  }
  for (auto &o : sk.always) { out.ket_cols.push_back(o.col); out.ket_sign.push_back(1.0); }
  for (size_t i = 0; i < sk.sometimes.size(); ++i) {
    out.ket_cols.push_back(sk.sometimes[i].col); out.ket_sign.push_back(sk.sometimes_sign[i]);
  }

  // occupation masks over the sometimes part.  Row t of `rest` is occupied if it is a filled orbital,
  // the physical orbital with p = 1, or an entangled mode whose bit in the Schmidt vector's mask says
  // so (left vectors: bit set, right vectors: bit clear).  Evaluated through byte lookup tables of the
  // (possibly complemented) mask: three loads per row instead of a loop over the orbitals.
  struct MaskMap {
    uint64_t constant = 0, phys = 0;
    std::vector<uint64_t> tab;   // 8 x 256
    int nbytes = 0;
    bool complement = false;
    uint64_t operator()(uint64_t m, int p) const {
      if (complement) m = ~m;
      uint64_t r = constant | (p == 1 ? phys : 0);
      for (int b = 0; b < nbytes; ++b) r |= tab[(size_t)b * 256 + ((m >> (8 * b)) & 255)];
      return r;
    }
  };
  auto make_map = [&](const std::vector<Entry> &rest) {
    MaskMap mm;
    mm.complement = (mode != 0);
    uint64_t bit_of[64];
    for (int i = 0; i < 64; ++i) bit_of[i] = 0;
    int top = -1;
    for (size_t t = 0; t < rest.size(); ++t) {
      const Orb &o = rest[t].o;
      if (o.kind == PHYS) mm.phys |= (1ull << t);
      else if (o.kind == FILLED) mm.constant |= (1ull << t);
      else { bit_of[o.idx] |= (1ull << t); top = std::max(top, o.idx); }
    }
    mm.nbytes = (top + 8) / 8;
    mm.tab.assign((size_t)std::max(mm.nbytes, 1) * 256, 0);
    for (int b = 0; b < mm.nbytes; ++b)
      for (int v = 1; v < 256; ++v) {
        const int low = __builtin_ctz(v);
        mm.tab[(size_t)b * 256 + v] = mm.tab[(size_t)b * 256 + (v & (v - 1))] | bit_of[8 * b + low];
      }
    return mm;
  };
  const MaskMap map_bra = make_map(br), map_ket = make_map(kr);
  const int qc = (mode == 0) ? 1 : -1;                       // slater.py:1111
  const int qtotal = (mode == 0) ? 0 : nferm_ket - nferm_bra;  // :1092
  const int n_rows = physical ? 2 * chi_bra : chi_bra;
  std::vector<int> rp(n_rows), ra(n_rows), rq(n_rows), ord(n_rows);
  int qmin = 1 << 30, qmax = -(1 << 30);
  for (int r = 0; r < n_rows; ++r) {
    rp[r] = physical ? r / chi_bra : -1;
    ra[r] = physical ? r % chi_bra : r;
    int p = physical ? rp[r] : 0;
    rq[r] = charge_bra[ra[r]] + (mode == 0 ? p : -p);  // charge left of the combined leg
    qmin = std::min(qmin, rq[r]);
    qmax = std::max(qmax, rq[r]);
  }
  {  // stable counting sort by the pipe charge (slater.py:1053-1058)
    std::vector<int> start((size_t)(qmax - qmin + 2), 0);
    for (int r = 0; r < n_rows; ++r) ++start[rq[r] - qmin + 1];
    for (size_t q = 1; q < start.size(); ++q) start[q] += start[q - 1];
    for (int r = 0; r < n_rows; ++r) ord[start[rq[r] - qmin]++] = r;
  }
  out.row_p.resize(n_rows); out.row_alpha.resize(n_rows); out.bra_masks.resize(n_rows);
  std::vector<int> q_sorted(n_rows);
  for (int r = 0; r < n_rows; ++r) {
    int s = ord[r];
    out.row_p[r] = rp[s];
    out.row_alpha[r] = ra[s];
    q_sorted[r] = rq[s];
    out.bra_masks[r] = map_bra(masks_bra[ra[s]], physical ? rp[s] : 0);
  }
  out.ket_masks.resize(chi_ket);
  for (int c = 0; c < chi_ket; ++c) out.ket_masks[c] = map_ket(masks_ket[c], 0);

  // charge blocks (slater.py:1132-1141)
  out.blocks.clear();
  int c0 = 0;
  while (c0 < chi_ket) {
    int c1 = c0;
    while (c1 < chi_ket && charge_ket[c1] == charge_ket[c0]) ++c1;
    const int q_bra = charge_ket[c0] + qtotal * qc;
    auto lo = std::lower_bound(q_sorted.begin(), q_sorted.end(), q_bra);
    auto hi = std::upper_bound(q_sorted.begin(), q_sorted.end(), q_bra);
    if (hi > lo) {
      const int r0 = (int)(lo - q_sorted.begin()), nr = (int)(hi - lo);
      const int n = __builtin_popcountll(out.ket_masks[c0]);
      for (int r = r0; r < r0 + nr; ++r)
        if (__builtin_popcountll(out.bra_masks[r]) != n)
          throw std::runtime_error("particle numbers of bra and ket block differ (slater.py:855)");
      for (int c = c0; c < c1; ++c)
        if (__builtin_popcountll(out.ket_masks[c]) != n)
          throw std::runtime_error("particle numbers within a ket block differ (slater.py:852)");
      int b[6] = {r0, nr, c0, c1 - c0, n, charge_ket[c0]};
      out.blocks.insert(out.blocks.end(), b, b + 6);
    }
    c0 = c1;
  }
  tmf_site_plan &h = out.h;
  h.mode = mode; h.physical = physical ? 1 : 0; h.n_bra = n_bra; h.n_ket = n_ket;
  h.k_bra = k_bra; h.k_ket = k_ket; h.f_bra = f_bra; h.f_ket = f_ket; h.k_always = k;
  h.s_bra = (int)br.size(); h.s_ket = (int)kr.size(); h.n_rows = n_rows;
  h.chi_bra = chi_bra; h.chi_ket = chi_ket; h.n_blocks = (int)out.blocks.size() / 6;
  h.qtotal = qtotal;
  h.ka_bra = kb; h.ka_ket = kk;
}

}  // namespace tmf

// ---------------------------------------------------------------------------------------------
// C ABI wrappers
// ---------------------------------------------------------------------------------------------
namespace tmf { const char *last_error_cstr(); }

extern "C" {

const char *tmf_last_error(void) { return tmf::last_error_cstr(); }

static tmf::TruncPar make_tp(int chi_max, double svd_min, double deg_tol, const int *sectors,
                             int n_sectors) {
  tmf::TruncPar tp;
  tp.chi_max = chi_max;
  tp.svd_min = svd_min;
  tp.degeneracy_tol = deg_tol;
  if (sectors != nullptr && n_sectors >= 0) {
    tp.filter = true;
    tp.sectors.assign(sectors, sectors + n_sectors);
  }
  return tp;
}

int tmf_lowest_sums(const double *a, int k, double base, int chi_max, double svd_min,
                    double degeneracy_tol, const int *sectors, int n_sectors, int filled_left,
                    int filled_right, int cap, double *sums_out, uint64_t *sets_out, int *n_out,
                    int *n_checked) {
  try {
    tmf::TruncPar tp = make_tp(chi_max, svd_min, degeneracy_tol, sectors, n_sectors);
    std::vector<double> sums;
    std::vector<uint64_t> sets;
    tmf::lowest_sums(a, k, base, tp, filled_left, filled_right, sums, sets, n_checked);
    if ((int)sums.size() > cap) {
      tmf::set_error("tmf_lowest_sums: output capacity exceeded");
      return TMF_ERR_VALUE;
    }
    std::copy(sums.begin(), sums.end(), sums_out);
    std::copy(sets.begin(), sets.end(), sets_out);
    *n_out = (int)sums.size();
    return TMF_OK;
  } catch (const std::invalid_argument &ex) {
    tmf::set_error(ex.what());
    return TMF_ERR_VALUE;
  } catch (const std::exception &ex) {
    tmf::set_error(ex.what());
    return TMF_ERR_RUNTIME;
  }
}

int tmf_bond_vectors_batched(int nbonds, const double *e, const int *k, const int *filled_left,
                             int chi_max, double svd_min, double degeneracy_tol,
                             const int *sectors, int n_sectors, int cap, uint64_t *masks,
                             double *lam, int *charge, int *chi, int *sec_q, int *sec_start,
                             int *sec_n, int n_threads) {
  tmf::TruncPar tp = make_tp(chi_max, svd_min, degeneracy_tol, sectors, n_sectors);
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = std::min(n_threads, std::max(1, nbonds));
  std::vector<int> status(n_threads, TMF_OK);
  std::vector<std::string> msgs(n_threads);
  const int S = TMF_MAX_MODES + 2;
  auto work = [&](int t) {
    for (int b = t; b < nbonds; b += n_threads) {
      try {
        tmf::BondVectors bv;
        tmf::bond_vectors(e + (size_t)b * TMF_MAX_MODES, k[b], filled_left[b], tp, bv);
        const int n = (int)bv.masks.size();
        if (n > cap) throw std::invalid_argument("bond vector capacity exceeded");
        std::copy(bv.masks.begin(), bv.masks.end(), masks + (size_t)b * cap);
        std::copy(bv.lam.begin(), bv.lam.end(), lam + (size_t)b * cap);
        std::copy(bv.charge.begin(), bv.charge.end(), charge + (size_t)b * cap);
        chi[b] = n;
        sec_n[b] = (int)bv.sec_q.size();
        std::copy(bv.sec_q.begin(), bv.sec_q.end(), sec_q + (size_t)b * S);
        std::copy(bv.sec_start.begin(), bv.sec_start.end(), sec_start + (size_t)b * S);
      } catch (const std::invalid_argument &ex) {
        status[t] = TMF_ERR_VALUE;
        msgs[t] = ex.what();
        return;
      } catch (const std::exception &ex) {
        status[t] = TMF_ERR_RUNTIME;
        msgs[t] = ex.what();
        return;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
  for (int t = 0; t < n_threads; ++t)
    if (status[t] != TMF_OK) {
      tmf::set_error(msgs[t]);
      return status[t];
    }
  return TMF_OK;
}

int tmf_slater_site_plan(int mode, int n_bra, int n_ket, int k_bra, int f_bra, int nferm_bra,
                         int chi_bra, const uint64_t *masks_bra, const int *charge_bra, int k_ket,
                         int f_ket, int nferm_ket, int chi_ket, const uint64_t *masks_ket,
                         const int *charge_ket, tmf_site_plan *plan, int *bra_cols,
                         double *bra_sign, int *ket_cols, double *ket_sign, uint64_t *bra_masks,
                         uint64_t *ket_masks, int *row_p, int *row_alpha, int *blocks) {
  try {
    tmf::SitePlan sp;
    tmf::site_plan(mode, n_bra, n_ket, k_bra, f_bra, nferm_bra, chi_bra, masks_bra, charge_bra,
                   k_ket, f_ket, nferm_ket, chi_ket, masks_ket, charge_ket, sp);
    *plan = sp.h;
    std::copy(sp.bra_cols.begin(), sp.bra_cols.end(), bra_cols);
    std::copy(sp.bra_sign.begin(), sp.bra_sign.end(), bra_sign);
    std::copy(sp.ket_cols.begin(), sp.ket_cols.end(), ket_cols);
    std::copy(sp.ket_sign.begin(), sp.ket_sign.end(), ket_sign);
    std::copy(sp.bra_masks.begin(), sp.bra_masks.end(), bra_masks);
    std::copy(sp.ket_masks.begin(), sp.ket_masks.end(), ket_masks);
    std::copy(sp.row_p.begin(), sp.row_p.end(), row_p);
    std::copy(sp.row_alpha.begin(), sp.row_alpha.end(), row_alpha);
    std::copy(sp.blocks.begin(), sp.blocks.end(), blocks);
    return TMF_OK;
  } catch (const std::invalid_argument &ex) {
    tmf::set_error(ex.what());
    return TMF_ERR_VALUE;
  } catch (const std::exception &ex) {
    tmf::set_error(ex.what());
    return TMF_ERR_ASSERT;
  }
}

}  // extern "C"
