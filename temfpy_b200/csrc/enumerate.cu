// K6/K7 on the device: best-first enumeration of the most probable Schmidt vectors of every bond.
//
// reference: schmidt_utils.lowest_sums (schmidt_utils.py:211-324: heap of (sum, push sequence number, i,
// subset), children "flip the next-larger |a|" / "swap to the next-larger |a|", loop while
// StoppingCondition.__call__ holds, then StoppingCondition.truncate :140-185) and
// SchmidtVectors.from_schmidt_modes (slater.py:633-700: left charges, stable sort by charge, sector table,
// Schmidt values :472-489).
//
// The search of one bond is inherently sequential (every pop decides the next pushes), but the bonds are
// independent: one warp per bond, lane 0 walks the search with a bucket queue in shared memory (the popped
// sums never decrease and only the window [base, base + max_logval] matters, so binning by sum gives the
// (sum, seq)-minimum after a scan of one or two items -- the pop order is exactly that of the reference's
// heap), then the whole warp truncates, sorts by charge (stable counting sort) and multiplies the Schmidt
// values.  A chain of 1025 bonds at chi = 1024 takes well under a millisecond of one wave of single warps,
// next to ~10 ms of 16 host threads; and it shards with the bonds over the GPUs, where the host cores do not.
//
// Everything that involves libm stays on the host so that the sums are bit-identical to the reference's:
// the weights a_i = log((1 - e_i) / e_i) / 2, their order, the base sum (NumPy's pairwise sum) and the
// thresholds.  The kernel only adds, subtracts, compares and multiplies (no contraction possible).
// A bond whose search does not fit the device structures (unbounded chi, window overflow) is flagged and
// recomputed by the host implementation (hostlogic.cpp).
#include <algorithm>
#include <cmath>
#include <functional>
#include <stdexcept>
#include <thread>

#include "cta.hpp"
#include "hostlogic.hpp"

namespace tmf {

constexpr int ENUM_NBK = 1024;          // buckets over the window of admissible sums
constexpr int ENUM_MAX_SECTORS = 66;    // charges span at most k + 1 <= 65 values

struct EnumJob {
  double magord[TMF_MAX_MODES];        // |a| in ascending order (k entries)
  double tab[2 * TMF_MAX_MODES];       // tab[2 i] = 1 - e_i, tab[2 i + 1] = e_i
  unsigned char order[TMF_MAX_MODES];  // mode index of magord[j]
  uint64_t neg;                        // modes with a_i < 0 (occupied in the most probable vector)
  double base;                         // sum of the negative weights
  uint64_t *masks;                     // outputs, `cap` entries each
  double *lam;
  int *charge;
  double *sums_scr;                    // scratch, cap + 2 entries each
  uint64_t *sets_scr;
  int *head;                           // [0] chi, [1] n_sectors, [2] status, [3] pops, then sec_q[66], sec_start[67]
  int k, filled_left;
};
static_assert(sizeof(EnumJob) % 8 == 0, "EnumJob must stay 8-byte aligned in arrays");

// status codes in head[2]
enum { ENUM_OK = 0, ENUM_FALLBACK = 1, ENUM_EMPTY = 2, ENUM_NOCUT = 3 };

struct EnumPar {
  int chi_max;            // >= 0 (unbounded searches run on the host)
  double max_logval;      // -log(svd_min) + degeneracy_tol            (schmidt_utils.py:96)
  double lim;             // -log(svd_min)                             (:171)
  double degeneracy_tol;
  const int *sectors;     // device array (nullptr: no filter)
  int n_sectors;
  int cap;                // capacity of the output arrays
  int pool;               // capacity of the live search frontier
};

TMF_DEVICE bool enum_is_sector(const EnumPar &p, int q) {
  if (p.sectors == nullptr) return true;
  for (int i = 0; i < p.n_sectors; ++i)
    if (p.sectors[i] == q) return true;
  return false;
}
TMF_DEVICE int enum_popc(uint64_t x) {
#if defined(TMF_HOSTSIM)
  return __builtin_popcountll(x);
#else
  return __popcll(x);
#endif
}

TMF_GLOBAL enumerate_kernel(const EnumJob *jobs, EnumPar par) {
  const EnumJob &jb = jobs[BLOCK_ID];
  const int k = jb.k;
  DYN_SMEM(unsigned char, raw);
  const int P = par.pool;
  double *psum = reinterpret_cast<double *>(raw);                       // P
  uint64_t *pset = reinterpret_cast<uint64_t *>(psum + P);              // P
  unsigned *pseq = reinterpret_cast<unsigned *>(pset + P);              // P
  int *cnt = reinterpret_cast<int *>(pseq + P);                         // ENUM_MAX_SECTORS + 2
  int *misc = cnt + ENUM_MAX_SECTORS + 2;                               // 8: n, cut, status
  unsigned *bits = reinterpret_cast<unsigned *>(misc + 8);              // ENUM_NBK / 32
  unsigned short *pnext = reinterpret_cast<unsigned short *>(bits + ENUM_NBK / 32);   // P
  unsigned short *head = pnext + P;                                     // ENUM_NBK
  unsigned char *pidx = reinterpret_cast<unsigned char *>(head + ENUM_NBK);           // P
  unsigned char *ord_s = pidx + ((P + 7) & ~7);                         // TMF_MAX_MODES
  double *mag_s = reinterpret_cast<double *>(ord_s + TMF_MAX_MODES);    // TMF_MAX_MODES
  const unsigned short NIL = 0xFFFF;
  PAR_FOR(i, TMF_MAX_MODES) { mag_s[i] = jb.magord[i]; ord_s[i] = jb.order[i]; }

  PAR_FOR(b, ENUM_NBK) head[b] = NIL;
  PAR_FOR(b, ENUM_NBK / 32) bits[b] = 0u;
  PAR_FOR(q, ENUM_MAX_SECTORS + 2) cnt[q] = 0;
  CTA_SYNC();

  // ---- the search (one thread) -----------------------------------------------------------------
  PAR_FOR(one, 1) {
    int n = 0, status = ENUM_OK, pops = 1;
    double front = 0.0, back = 0.0;
    if (k == 0) {   // schmidt_utils.py:268-271
      if (enum_is_sector(par, jb.filled_left)) { jb.sums_scr[0] = 0.0; jb.sets_scr[0] = 0; n = 1; }
    } else {
      const double base = jb.base;
      const double scale = ENUM_NBK / (par.max_logval * 1.0001);
      if (enum_is_sector(par, jb.filled_left + enum_popc(jb.neg))) {   // :277-279
        jb.sums_scr[0] = base; jb.sets_scr[0] = jb.neg; n = 1;
        front = back = base;
      }
      // free list of pool slots
      for (int i = 0; i < P; ++i) pnext[i] = (unsigned short)(i + 1 < P ? i + 1 : NIL);
      unsigned short free_head = 0;
      int cur = ENUM_NBK;
      // single best item beyond the window (its successors would need a heap: fall back to the host)
      bool have_over = false;
      int n_over = 0;
      double o_sum = 0.0; uint64_t o_set = 0; unsigned o_seq = 0; int o_i = 0;
      unsigned seq = 0;
      auto push = [&](double s, unsigned sq, int ii, uint64_t set) {
        const double t = (s - base) * scale;
        if (!(t < (double)ENUM_NBK)) {
          ++n_over;
          if (!have_over || s < o_sum || (s == o_sum && sq < o_seq)) { o_sum = s; o_set = set; o_seq = sq; o_i = ii; }
          have_over = true;
          return;
        }
        if (free_head == NIL) { status = ENUM_FALLBACK; return; }
        const int b = t > 0.0 ? (int)t : 0;
        const unsigned short id = free_head;
        free_head = pnext[id];
        psum[id] = s; pset[id] = set; pseq[id] = sq; pidx[id] = (unsigned char)ii;
        pnext[id] = head[b];
        head[b] = id;
        bits[b >> 5] |= (1u << (b & 31));
        if (b < cur) cur = b;
      };
      push(base + mag_s[0], seq, 0, jb.neg ^ (1ull << ord_s[0]));   // :291-293
      for (;;) {
        if (status != ENUM_OK) break;
        // more_needed (StoppingCondition.__call__, schmidt_utils.py:99-138) on the sums collected so far
        if (n > 0 && (n > par.chi_max || back - front > par.max_logval)) break;
        // lowest non-empty bucket
        if (!(cur < ENUM_NBK && head[cur] != NIL)) {
          int w = cur >> 5;
          bool found = false;
          if (cur < ENUM_NBK) {
            unsigned m = bits[w] & (~0u << (cur & 31));
            for (;;) {
              if (m) {
#if defined(TMF_HOSTSIM)
                cur = (w << 5) + __builtin_ctz(m);
#else
                cur = (w << 5) + (__ffs((int)m) - 1);
#endif
                found = true;
                break;
              }
              if (++w >= ENUM_NBK / 32) break;
              m = bits[w];
            }
          }
          if (!found) cur = ENUM_NBK;
        }
        double s; uint64_t set; int ii;
        if (cur < ENUM_NBK) {
          unsigned short best = head[cur], bprev = NIL, prev = head[cur];
          for (unsigned short it = pnext[prev]; it != NIL; prev = it, it = pnext[it])
            if (psum[it] < psum[best] || (psum[it] == psum[best] && pseq[it] < pseq[best])) { best = it; bprev = prev; }
          if (bprev == NIL) head[cur] = pnext[best]; else pnext[bprev] = pnext[best];
          if (head[cur] == NIL) bits[cur >> 5] &= ~(1u << (cur & 31));
          s = psum[best]; set = pset[best]; ii = pidx[best];
          pnext[best] = free_head;
          free_head = best;
        } else if (have_over) {
          // The best item beyond the window.  Popping it normally ends the search (its sum exceeds the admissible
          // range); if the search had to go on (sector filters), the order of the other overflow items would be
          // needed -- they were not kept: fall back to the host.
          s = o_sum; set = o_set; ii = o_i;
          if (n_over > 1) {
            const bool accepted = enum_is_sector(par, jb.filled_left + enum_popc(set));
            const double fr = (n == 0) ? s : front;
            const bool ends = accepted && (n + 1 > par.chi_max || s - fr > par.max_logval);
            if (!ends) { status = ENUM_FALLBACK; break; }
          }
          have_over = false;
          n_over = 0;
        } else {
          break;   // queue empty
        }
        ++pops;
        if (enum_is_sector(par, jb.filled_left + enum_popc(set))) {
          if (n >= par.cap + 1) { status = ENUM_FALLBACK; break; }
          jb.sums_scr[n] = s; jb.sets_scr[n] = set;
          if (n == 0) front = s;
          back = s;
          ++n;
        }
        if (ii < k - 1) {   // :304-315
          const uint64_t c1 = set ^ (1ull << ord_s[ii + 1]);
          double s1 = s + mag_s[ii + 1];
          push(s1, ++seq, ii + 1, c1);
          const uint64_t c2 = c1 ^ (1ull << ord_s[ii]);
          s1 = s1 - mag_s[ii];
          push(s1, ++seq, ii + 1, c2);
        }
      }
    }
    if (status == ENUM_OK && n == 0) status = ENUM_EMPTY;
    misc[0] = n;
    misc[1] = -1;
    misc[2] = status;
    misc[3] = pops;
  }
  CTA_SYNC();
  const int n = misc[0];
  if (misc[2] != ENUM_OK) {
    PAR_FOR(one, 1) { jb.head[0] = 0; jb.head[1] = 0; jb.head[2] = misc[2]; jb.head[3] = misc[3]; }
    return;
  }
  // ---- truncate (schmidt_utils.py:140-185): last index that satisfies all three conditions ----------
  {
    const double *lv = jb.sums_scr;
    const double lv0 = lv[0];
    PAR_FOR(t, NTHREADS) {
      int best = -1;
      for (int i = t; i < n; i += NTHREADS) {
        bool ok = !(i >= par.chi_max) && (lv[i] - lv0 < par.lim);
        if (ok && i < n - 1 && !((lv[i + 1] - lv[i]) > par.degeneracy_tol)) ok = false;
        if (ok) best = i;
      }
      reinterpret_cast<int *>(psum)[t] = best;     // the pool is dead by now
    }
    CTA_SYNC();
    PAR_FOR(one, 1) {
      int best = -1;
      for (int t = 0; t < NTHREADS; ++t) best = best > reinterpret_cast<int *>(psum)[t] ? best : reinterpret_cast<int *>(psum)[t];
      misc[1] = best;
    }
    CTA_SYNC();
  }
  const int chi = misc[1] + 1;
  if (chi <= 0 || chi > par.cap) {
    PAR_FOR(one, 1) { jb.head[0] = 0; jb.head[1] = 0; jb.head[2] = chi <= 0 ? ENUM_NOCUT : ENUM_FALLBACK; jb.head[3] = misc[3]; }
    return;
  }
  // ---- stable counting sort by the left charge (slater.py:673-683) ----------------------------------
  // rank[i] (position in the sorted order) is kept in pseq (the pool is dead)
  unsigned *rank = pseq;
#if defined(TMF_HOSTSIM)
  PAR_FOR(one, 1) {
    for (int i = 0; i < chi; ++i) ++cnt[enum_popc(jb.sets_scr[i]) + 1];
    for (int q = 1; q <= ENUM_MAX_SECTORS; ++q) cnt[q] += cnt[q - 1];
    for (int i = 0; i < chi; ++i) rank[i] = (unsigned)cnt[enum_popc(jb.sets_scr[i])]++;
    // cnt[q] is now the END of sector q; rebuild the starts below from the ends
  }
  CTA_SYNC();
#else
  {
    // one warp: chunks of 32 entries in index order; inside a chunk the lanes with equal charge are ranked
    // by lane number (match_any), across chunks by the running counters -> stable
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
      for (int i0 = 0; i0 < chi; i0 += 32) {
        const int i = i0 + lane;
        const int q = (i < chi) ? enum_popc(jb.sets_scr[i]) : ENUM_MAX_SECTORS;
        const unsigned peers = __match_any_sync(0xffffffffu, q);
        if (i < chi && lane == (__ffs((int)peers) - 1)) cnt[q + 1] += __popc(peers);
      }
      __syncwarp();
      if (lane == 0)
        for (int q = 1; q <= ENUM_MAX_SECTORS; ++q) cnt[q] += cnt[q - 1];
      __syncwarp();
      for (int i0 = 0; i0 < chi; i0 += 32) {
        const int i = i0 + lane;
        const int q = (i < chi) ? enum_popc(jb.sets_scr[i]) : ENUM_MAX_SECTORS;
        const unsigned peers = __match_any_sync(0xffffffffu, q);
        if (i < chi) rank[i] = (unsigned)(cnt[q] + __popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
        if (i < chi && lane == (__ffs((int)peers) - 1)) cnt[q] += __popc(peers);
        __syncwarp();
      }
    }
    __syncthreads();
  }
#endif
  // ---- outputs: masks, charges, Schmidt values (slater.py:489: product in mode order) ----------------
  PAR_FOR(i, chi) {
    const uint64_t m = jb.sets_scr[i];
    const int r = (int)rank[i];
    jb.masks[r] = m;
    jb.charge[r] = jb.filled_left + enum_popc(m);
    double p = 1.0;
    for (int j = 0; j < k; ++j) p *= jb.tab[2 * j + (int)((m >> j) & 1)];
    jb.lam[r] = sqrt(p);
  }
  // sector table: after the ranking pass cnt[q] is the end of sector q (q = number of occupied modes)
  PAR_FOR(one, 1) {
    int ns = 0, start = 0;
    int *sec_q = jb.head + 4, *sec_start = jb.head + 4 + ENUM_MAX_SECTORS;
    for (int q = 0; q < ENUM_MAX_SECTORS; ++q) {
      const int end = cnt[q];
      if (end > start) {
        sec_q[ns] = jb.filled_left + q;
        sec_start[ns] = start;
        ++ns;
        start = end;
      }
    }
    sec_start[ns] = chi;
    jb.head[0] = chi; jb.head[1] = ns; jb.head[2] = ENUM_OK; jb.head[3] = misc[3];
  }
}

static size_t enum_smem_bytes(int pool) {
  return (size_t)pool * (8 + 8 + 4 + 2 + 1) + sizeof(int) * (ENUM_MAX_SECTORS + 2 + 8) + ENUM_NBK / 8 + 2 * ENUM_NBK +
         9 * TMF_MAX_MODES + 128;
}

// Host driver: prepares the per-bond inputs (everything that needs libm), runs the kernel, brings the tables
// back and fills `out[b]`.  Bonds the device could not finish are recomputed by bond_vectors().
// e: nb x TMF_MAX_MODES left eigenvalues; work_dev: tmf_enum_workspace(nb, chi_max) bytes.
int64_t enum_workspace_bytes(int nb, int chi_max) {
  if (chi_max < 0) return 256;
  const int64_t cap = chi_max + 2;
  return align256((int64_t)nb * sizeof(EnumJob)) + 5 * align256((int64_t)nb * (cap + 2) * 8) +
         align256((int64_t)nb * 4 * (4 + 2 * ENUM_MAX_SECTORS + 2)) + 4096;
}

int enumerate_device(int nb, const double *const *e_ptr, const int *k, const int *filled_left, const TruncPar &tp,
                     std::vector<BondVectors *> &out, void *work_dev, int64_t work_bytes, void *stream,
                     unsigned char *stage, size_t stage_bytes, int n_threads, EnumResident *keep) {
  if (keep) keep->resident = false;
  const double max_logval = -std::log(tp.svd_min) + tp.degeneracy_tol;
  const int pool = (std::min(60000, std::max(tp.chi_max, 0) + 96) + 7) & ~7;   // multiple of 8: keeps the shared arrays aligned
  const bool device_ok = tp.chi_max >= 0 && max_logval > 0.0 && max_logval < 64.0 && pool < 65000 &&
                         enum_smem_bytes(pool) <= 200 * 1024 && work_dev != nullptr &&
                         work_bytes >= enum_workspace_bytes(nb, tp.chi_max) && (!tp.filter || tp.sectors.size() <= 512);
  auto host_one = [&](int b) { bond_vectors(e_ptr[b], k[b], filled_left[b], tp, *out[b]); };
  if (!device_ok || nb == 0) {
    for (int b = 0; b < nb; ++b) host_one(b);
    return TMF_OK;
  }
  const int64_t cap = tp.chi_max + 2;
  Arena ar(work_dev, work_bytes);
  EnumJob *jobs_dev = ar.take<EnumJob>(nb);
  uint64_t *masks_dev = ar.take<uint64_t>((int64_t)nb * cap);
  double *lam_dev = ar.take<double>((int64_t)nb * cap);
  int *charge_dev = ar.take<int>((int64_t)nb * cap);
  double *sums_dev = ar.take<double>((int64_t)nb * (cap + 2));
  uint64_t *sets_dev = ar.take<uint64_t>((int64_t)nb * (cap + 2));
  const int HW = 4 + 2 * ENUM_MAX_SECTORS + 2;
  int *head_dev = ar.take<int>((int64_t)nb * HW);
  int *sectors_dev = ar.take<int>(tp.filter ? (int64_t)tp.sectors.size() + 1 : 1);
  if (!ar.ok()) { set_error("enumerate: workspace overflow"); return TMF_ERR_VALUE; }
  // pinned staging: jobs up, tables down
  const size_t up_bytes = (size_t)nb * sizeof(EnumJob);
  const size_t o_head = (up_bytes + 255) & ~size_t(255);
  const size_t o_masks = o_head + (((size_t)nb * HW * 4 + 255) & ~size_t(255));
  const size_t o_lam = o_masks + (size_t)nb * cap * 8, o_charge = o_lam + (size_t)nb * cap * 8;
  const size_t need = o_charge + (size_t)nb * cap * 4;
  std::vector<unsigned char> own;
  if (stage == nullptr || stage_bytes < need) { own.resize(need); stage = own.data(); }
  EnumJob *jobs = reinterpret_cast<EnumJob *>(stage);
  if (n_threads <= 0) n_threads = 4;
  n_threads = std::max(1, std::min(n_threads, nb));
  auto run_threads = [&](const std::function<void(int)> &f) { pool_for(nb, n_threads, f); };
  run_threads([&](int b) {
    EnumJob &j = jobs[b];
    std::memset(&j, 0, sizeof(j));
    const int kk = k[b];
    const double *e = e_ptr[b];
    std::vector<double> a(kk), negs;
    for (int i = 0; i < kk; ++i) a[i] = std::log((1.0 - e[i]) / e[i]) / 2;   // slater.py:428, :663
    if (tp.snap) snap_degenerate(a.data(), e, kk);
    for (int i = 0; i < kk; ++i)
      if (a[i] < 0) { negs.push_back(a[i]); j.neg |= (1ull << i); }
    // NumPy's pairwise sum of the negative weights (schmidt_utils.py:274), as in bond_vectors()
    {
      const int nn = (int)negs.size();
      double res;
      if (nn < 8) {
        res = 0.0;
        for (int i = 0; i < nn; ++i) res += negs[i];
      } else {
        double r[8];
        for (int t = 0; t < 8; ++t) r[t] = negs[t];
        int i = 8;
        for (; i < nn - (nn % 8); i += 8)
          for (int t = 0; t < 8; ++t) r[t] += negs[i + t];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < nn; ++i) res += negs[i];
      }
      j.base = res;
    }
    std::vector<int> order(kk);
    for (int i = 0; i < kk; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return std::fabs(a[x]) < std::fabs(a[y]); });
    for (int i = 0; i < kk; ++i) {
      j.magord[i] = std::fabs(a[order[i]]);
      j.order[i] = (unsigned char)order[i];
      j.tab[2 * i] = 1.0 - e[i];
      j.tab[2 * i + 1] = e[i];
    }
    j.k = kk;
    j.filled_left = filled_left[b];
    j.masks = masks_dev + (int64_t)b * cap;
    j.lam = lam_dev + (int64_t)b * cap;
    j.charge = charge_dev + (int64_t)b * cap;
    j.sums_scr = sums_dev + (int64_t)b * (cap + 2);
    j.sets_scr = sets_dev + (int64_t)b * (cap + 2);
    j.head = head_dev + (int64_t)b * HW;
  });
  int rc = copy_h2d(jobs_dev, jobs, up_bytes, stream);
  if (rc) return rc;
  EnumPar par;
  par.chi_max = tp.chi_max;
  par.max_logval = max_logval;
  par.lim = -std::log(tp.svd_min);
  par.degeneracy_tol = tp.degeneracy_tol;
  par.sectors = nullptr;
  par.n_sectors = 0;
  if (tp.filter) {
    rc = copy_h2d(sectors_dev, tp.sectors.data(), sizeof(int) * tp.sectors.size(), stream);
    if (rc) return rc;
    par.sectors = sectors_dev;
    par.n_sectors = (int)tp.sectors.size();
  }
  par.cap = (int)cap;
  par.pool = pool;
  rc = launch_t("enumerate", enumerate_kernel, nb, 32, enum_smem_bytes(pool), stream, (const EnumJob *)jobs_dev, par);
  if (rc) return rc;
  int *head = reinterpret_cast<int *>(stage + o_head);
  if (keep != nullptr) {
    // Resident form: only the heads (chi, sector tables, status) come to the host now; the tables stay in
    // device memory for the planning / tensor kernels and are downloaded later, off the critical path
    // (enumerate_fetch_tables).  A bond the device could not finish is recomputed on the host and uploaded.
    rc = copy_d2h_sync(head, head_dev, (size_t)nb * HW * 4, stream);
    if (rc) return rc;
    keep->resident = true;
    keep->nb = nb; keep->cap = cap; keep->hw = HW;
    keep->masks_dev = masks_dev; keep->lam_dev = lam_dev; keep->charge_dev = charge_dev; keep->head_dev = head_dev;
    keep->head_host.assign(head, head + (size_t)nb * HW);
    std::vector<int> fb;
    for (int b = 0; b < nb; ++b) {
      const int *h = keep->head_host.data() + (size_t)b * HW;
      if (h[2] == ENUM_EMPTY) throw std::runtime_error("-1|No Schmidt vectors left after filtering by `trunc_par.sectors`!");
      if (h[2] == ENUM_NOCUT) throw std::runtime_error("-2|truncate: no admissible cut");
      if (h[2] == ENUM_FALLBACK) fb.push_back(b);
    }
    if (!fb.empty()) {
      std::vector<BondVectors> bvs(fb.size());
      pool_for((int)fb.size(), n_threads, [&](int t) { bond_vectors(e_ptr[fb[t]], k[fb[t]], filled_left[fb[t]], tp, bvs[t]); });
      for (size_t t = 0; t < fb.size() && keep->resident; ++t) {
        const int b = fb[t];
        const BondVectors &bv = bvs[t];
        int *h = keep->head_host.data() + (size_t)b * HW;
        const int chi = (int)bv.masks.size(), ns = (int)bv.sec_q.size();
        if (chi > cap || ns > ENUM_MAX_SECTORS) { keep->resident = false; break; }   // does not fit: host path for everything
        h[0] = chi; h[1] = ns; h[2] = ENUM_OK;
        std::copy(bv.sec_q.begin(), bv.sec_q.end(), h + 4);
        std::copy(bv.sec_start.begin(), bv.sec_start.end(), h + 4 + ENUM_MAX_SECTORS);
        if ((rc = copy_h2d(head_dev + (size_t)b * HW, h, sizeof(int) * HW, stream))) return rc;
        if ((rc = copy_h2d(masks_dev + (size_t)b * cap, bv.masks.data(), sizeof(uint64_t) * chi, stream))) return rc;
        if ((rc = copy_h2d(lam_dev + (size_t)b * cap, bv.lam.data(), sizeof(double) * chi, stream))) return rc;
        if ((rc = copy_h2d(charge_dev + (size_t)b * cap, bv.charge.data(), sizeof(int) * chi, stream))) return rc;
      }
    }
    if (keep->resident) return TMF_OK;
    // (a table did not fit its slot: the classic path below re-reads everything and unpacks on the host)
  }
  rc = copy_d2h_async(head, head_dev, (size_t)nb * HW * 4, stream);
  if (rc) return rc;
  rc = copy_d2h_async(stage + o_masks, masks_dev, (size_t)nb * cap * 8, stream);
  if (rc) return rc;
  rc = copy_d2h_async(stage + o_lam, lam_dev, (size_t)nb * cap * 8, stream);
  if (rc) return rc;
  rc = copy_d2h_sync(stage + o_charge, charge_dev, (size_t)nb * cap * 4, stream);
  if (rc) return rc;
  const uint64_t *masks_h = reinterpret_cast<const uint64_t *>(stage + o_masks);
  const double *lam_h = reinterpret_cast<const double *>(stage + o_lam);
  const int *charge_h = reinterpret_cast<const int *>(stage + o_charge);
  std::vector<int> err(nb, 0);
  auto fill = [&](int b) {
    const int *h = head + (size_t)b * HW;
    BondVectors &o = *out[b];
    if (h[2] == ENUM_FALLBACK) { host_one(b); return; }
    if (h[2] == ENUM_EMPTY) { err[b] = 1; return; }
    if (h[2] == ENUM_NOCUT) { err[b] = 2; return; }
    const int chi = h[0], ns = h[1];
    o.k = k[b];
    o.filled_left = filled_left[b];
    o.masks.assign(masks_h + (size_t)b * cap, masks_h + (size_t)b * cap + chi);
    o.lam.assign(lam_h + (size_t)b * cap, lam_h + (size_t)b * cap + chi);
    o.charge.assign(charge_h + (size_t)b * cap, charge_h + (size_t)b * cap + chi);
    o.sec_q.assign(h + 4, h + 4 + ns);
    o.sec_start.assign(h + 4 + ENUM_MAX_SECTORS, h + 4 + ENUM_MAX_SECTORS + ns + 1);
  };
  run_threads(fill);
  for (int b = 0; b < nb; ++b) {
    if (err[b] == 1) throw std::runtime_error("-1|No Schmidt vectors left after filtering by `trunc_par.sectors`!");   // ValueError
    if (err[b] == 2) throw std::runtime_error("-2|truncate: no admissible cut");
  }
  return TMF_OK;
}

// Downloads the tables of a resident enumeration (asynchronously, into pinned `stage`) -- the caller
// synchronises (event / stream) and then unpacks with enumerate_unpack_tables.
size_t enumerate_tables_stage_bytes(const EnumResident &r) {
  return (size_t)r.nb * r.cap * (8 + 8 + 4) + 1024;
}
int enumerate_fetch_tables(const EnumResident &r, unsigned char *stage, void *stream) {
  const size_t o_lam = (size_t)r.nb * r.cap * 8, o_charge = 2 * o_lam;
  int rc = copy_d2h_async(stage, r.masks_dev, (size_t)r.nb * r.cap * 8, stream);
  if (rc) return rc;
  rc = copy_d2h_async(stage + o_lam, r.lam_dev, (size_t)r.nb * r.cap * 8, stream);
  if (rc) return rc;
  return copy_d2h_async(stage + o_charge, r.charge_dev, (size_t)r.nb * r.cap * 4, stream);
}
void enumerate_unpack_tables(const EnumResident &r, const unsigned char *stage, const int *k, const int *filled_left,
                             std::vector<BondVectors *> &out, int n_threads) {
  const size_t o_lam = (size_t)r.nb * r.cap * 8, o_charge = 2 * o_lam;
  const uint64_t *masks_h = reinterpret_cast<const uint64_t *>(stage);
  const double *lam_h = reinterpret_cast<const double *>(stage + o_lam);
  const int *charge_h = reinterpret_cast<const int *>(stage + o_charge);
  pool_for(r.nb, std::max(1, n_threads), [&](int b) {
    const int *h = r.head_host.data() + (size_t)b * r.hw;
    BondVectors &o = *out[b];
    const int chi = h[0], ns = h[1];
    o.k = k[b];
    o.filled_left = filled_left[b];
    o.masks.assign(masks_h + (size_t)b * r.cap, masks_h + (size_t)b * r.cap + chi);
    o.lam.assign(lam_h + (size_t)b * r.cap, lam_h + (size_t)b * r.cap + chi);
    o.charge.assign(charge_h + (size_t)b * r.cap, charge_h + (size_t)b * r.cap + chi);
    o.sec_q.assign(h + 4, h + 4 + ns);
    o.sec_start.assign(h + 4 + ENUM_MAX_SECTORS, h + 4 + ENUM_MAX_SECTORS + ns + 1);
  });
}

}  // namespace tmf
