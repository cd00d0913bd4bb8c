// Pfaffian (Bogoliubov) path: kernels that exist only for pfaffian.py.
//
// The Nambu correlation matrix in the Majorana basis is C_M = 1/2 + iA (A real antisymmetric,
// pfaffian.py:269-273).  Its real representation P' (re/im interleaved, 4L x 4L) is a real symmetric
// projector of rank 2L, so the per-bond eigenproblems of pfaffian.py:789 run through the *same*
// mode-extraction kernels as the Slater path (tmf_slater_modes_batched on P', cuts at 4x): every
// complex eigenvector w of a Majorana block shows up as a two-dimensional real eigenspace
// span{emb(w), J emb(w)} (J = multiplication by i).
//
//  K4p  pair_modes_kernel   picks one complex mode per J-invariant plane (pivoted Gram-Schmidt inside
//       windows of numerically degenerate eigenvalues) for the eigenvalues e <= 1/2 and writes, per
//       mode a (e ascending, the order of the reference's `e`, pfaffian.py:839/846), four real
//       columns 4a..4a+3 = emb(w), J emb(w), emb(conj w), J emb(conj w); conj w is the Nambu partner
//       (eigenvalue 1-e) that pfaffian.py:886/891 constructs.
//  K11  pfaffians_kernel    all Pf(N[idx, idx]) of a (bra sector, ket sector) block, complex128,
//       Parlett-Reid with partial pivoting -- replaces the pfapack call of pfaffian.py:1425 and the
//       gather of :1468-1477.  One warp per tensor entry, the m x m sub-matrix lives in a private
//       shared-memory tile of the warp (__syncwarp only), N itself is staged once per CTA.
#include "cta.hpp"

namespace tmf {

#if defined(TMF_HOSTSIM)
#define PF_LANE_FOR(l) for (int l = 0; l < 32; ++l)
#define PF_WARP_FOR(w, W) for (int w = 0; w < (W); ++w)
#define PF_WSYNC() ((void)0)
#else
#define PF_LANE_FOR(l) for (int l = (threadIdx.x & 31), l##_once = 1; l##_once; l##_once = 0)
#define PF_WARP_FOR(w, W) for (int w = (threadIdx.x >> 5), w##_once = 1; w##_once && w < (W); w##_once = 0)
#define PF_WSYNC() __syncwarp()
#endif

static_assert(sizeof(tmf_pair_job) == 64, "pair descriptor must be 64 bytes");
static_assert(sizeof(tmf_pf_block) == 64, "pfaffian block descriptor must be 64 bytes");

constexpr int PAIR_MAXC = 64;   // columns of a degenerate window (4k <= TMF_MAX_MODES)

// (J v)[2q] = -v[2q+1], (J v)[2q+1] = v[2q]
TMF_DEVICE double jcomp(const double *v, int r) { return (r & 1) ? v[r - 1] : -v[r + 1]; }

// Pivoted Gram-Schmidt step on the window `win` (W columns of length rows): picks the column of
// largest remaining norm, normalises it into ubuf and removes its component (and, if `with_j`, the
// component along J ubuf) from every column of the window.  Returns the squared norm of the pick.
TMF_DEVICE double gs_step(double *win, int W, int rows, double *ubuf, double *part, double *ca, double *cb,
                          double *red, bool with_j) {
  PAR_FOR(item, W * 32) {
    const int c = item >> 5, lane = item & 31;
    const double *y = win + (int64_t)c * rows;
    double s = 0.0;
    for (int r = lane; r < rows; r += 32) s += y[r] * y[r];
    part[c * 33 + lane] = s;
  }
  CTA_SYNC();
  PAR_FOR(one, 1) {
    double best = -1.0;
    int p = 0;
    for (int c = 0; c < W; ++c) {
      double s = 0.0;
      for (int l = 0; l < 32; ++l) s += part[c * 33 + l];
      if (s > best) { best = s; p = c; }
    }
    red[0] = (double)p;
    red[1] = (best > 0.0) ? 1.0 / sqrt(best) : 0.0;
    red[2] = best;
  }
  CTA_SYNC();
  const int p = (int)red[0];
  const double inv = red[1], best = red[2];
  PAR_FOR(r, rows) ubuf[r] = win[(int64_t)p * rows + r] * inv;
  CTA_SYNC();
  PAR_FOR(item, W * 32) {
    const int c = item >> 5, lane = item & 31;
    const double *y = win + (int64_t)c * rows;
    double sa = 0.0, sb = 0.0;
    for (int r = lane; r < rows; r += 32) {
      sa += ubuf[r] * y[r];
      if (with_j) sb += jcomp(ubuf, r) * y[r];
    }
    part[c * 33 + lane] = sa;
    part[(PAIR_MAXC + c) * 33 + lane] = sb;
  }
  CTA_SYNC();
  PAR_FOR(c, W) {
    double sa = 0.0, sb = 0.0;
    for (int l = 0; l < 32; ++l) {
      sa += part[c * 33 + l];
      sb += part[(PAIR_MAXC + c) * 33 + l];
    }
    ca[c] = sa;
    cb[c] = sb;
  }
  CTA_SYNC();
  PAR_FOR(idx, W * rows) {
    const int c = idx / rows, r = idx - c * rows;
    win[(int64_t)c * rows + r] -= ca[c] * ubuf[r] + (with_j ? cb[c] * jcomp(ubuf, r) : 0.0);
  }
  CTA_SYNC();
  return best;
}

// writes the four real columns of complex mode `a` (emb(w) in ubuf) into the staging buffer
TMF_DEVICE void emit_mode(double *outb, int a, int rows, const double *ubuf) {
  PAR_FOR(r, rows) {
    const double u = ubuf[r], ju = jcomp(ubuf, r);
    double *o = outb + (int64_t)(4 * a) * rows;
    o[r] = u;                                   // emb(w)
    o[rows + r] = ju;                           // J emb(w)
    o[2 * rows + r] = (r & 1) ? -u : u;         // emb(conj w)
    o[3 * rows + r] = (r & 1) ? ju : -ju;       // J emb(conj w): [2q] = u[2q+1], [2q+1] = u[2q]
  }
}

TMF_GLOBAL pair_modes_kernel(const tmf_pair_job *jobs, double win_rel, double win_abs, double half_tol) {
  const tmf_pair_job jb = jobs[BLOCK_ID];
  const int rows = jb.rows, k4 = jb.k4, k = k4 / 4;
  DYN_SMEM(double, sm);
  double *ec = sm;                         // PAIR_MAXC   block eigenvalue of candidate t (ascending)
  double *part = ec + PAIR_MAXC;           // 2 * PAIR_MAXC * 33
  double *ca = part + 2 * PAIR_MAXC * 33;  // PAIR_MAXC
  double *cb = ca + PAIR_MAXC;             // PAIR_MAXC
  double *red = cb + PAIR_MAXC;            // 8
  if (k4 <= 0 || rows <= 0) {
    PAR_FOR(one, 1) { *jb.status = 0; *jb.kh_out = 0; }
    return;
  }
  if ((k4 & 3) || k4 > PAIR_MAXC) {
    PAR_FOR(one, 1) *jb.status = 2;
    return;
  }
  double *win = jb.tmp;                              // rows x PAIR_MAXC: current window
  double *ubuf = win + (int64_t)PAIR_MAXC * rows;    // rows: accepted vector
  double *rbuf = ubuf + rows;                        // rows x PAIR_MAXC / 2: real basis of the 1/2 space
  double *outb = rbuf + (int64_t)(PAIR_MAXC / 2) * rows;  // rows x k4: staged result
  // candidate t (ascending block eigenvalue) -> raw column
  //   side L: raw columns are ordered by decreasing block eigenvalue -> reversed
  //   side R: raw columns are ordered by increasing block eigenvalue
  PAR_FOR(t, k4) ec[t] = (jb.side == TMF_SIDE_L) ? jb.e_raw[k4 - 1 - t] : 1.0 - jb.e_raw[t];
  PAR_FOR(one, 1) red[4] = 0.0;   // status
  CTA_SYNC();
  int nh = 0;
  for (int t = 0; t < k4; ++t) nh += (fabs(ec[t] - 0.5) <= half_tol) ? 1 : 0;
  if (nh & 3) {
    PAR_FOR(one, 1) *jb.status = 5;   // 1/2 eigenvalues asymmetrical in spectrum (pfaffian.py:805)
    return;
  }
  const int kh = nh / 4, k2 = 2 * (k - kh);   // k2 regular candidates (e < 1/2), then nh half columns
  int acc = 0, i = 0;
  while (i < k2) {
    int j = i + 1;
    while (j < k2 && fabs(ec[j] - ec[j - 1]) <= win_rel * fabs(ec[j]) + win_abs) ++j;
    const int W = j - i, m = W / 2;
    if (W & 1) {
      PAR_FOR(one, 1) red[4] = 3.0;   // odd multiplicity: not a J-invariant eigenspace
    }
    PAR_FOR(idx, W * rows) {
      const int c = idx / rows, r = idx - c * rows;
      const int raw = (jb.side == TMF_SIDE_L) ? (k4 - 1 - (i + c)) : (i + c);
      win[(int64_t)c * rows + r] = jb.V[(int64_t)raw * jb.ld + r];
    }
    CTA_SYNC();
    for (int t = 0; t < m; ++t) {
      const double best = gs_step(win, W, rows, ubuf, part, ca, cb, red, true);
      if (best < 0.25) {
        PAR_FOR(one, 1) red[4] = 4.0;   // the plane was not there: eigenvectors are inconsistent
      }
      emit_mode(outb, acc, rows, ubuf);
      PAR_FOR(one, 1) jb.e_out[acc] = ec[i + 2 * t];
      CTA_SYNC();
      ++acc;
    }
    i = j;
  }
  if (kh > 0) {
    // eigenvalue 1/2 (pfaffian.py:802-816): the eigenspace is the complexification of the real null
    // space of the block of A; its real vectors have vanishing odd (imaginary) components.
    const int W = nh;
    PAR_FOR(idx, W * rows) {
      const int c = idx / rows, r = idx - c * rows;
      const int t = k2 + c;
      const int raw = (jb.side == TMF_SIDE_L) ? (k4 - 1 - t) : t;
      win[(int64_t)c * rows + r] = (r & 1) ? 0.0 : jb.V[(int64_t)raw * jb.ld + r];
    }
    CTA_SYNC();
    for (int t = 0; t < 2 * kh; ++t) {
      const double best = gs_step(win, W, rows, ubuf, part, ca, cb, red, false);
      if (best < 1e-3) {
        PAR_FOR(one, 1) red[4] = 6.0;   // 1/2 eigenvectors cannot be made real (pfaffian.py:813)
      }
      PAR_FOR(r, rows) rbuf[(int64_t)t * rows + r] = ubuf[r];
      CTA_SYNC();
    }
    // complex pairs w_j = (r_j + i r_{kh+j}) / sqrt(2)  (pfaffian.py:884)
    for (int j = 0; j < kh; ++j) {
      PAR_FOR(r, rows) {
        const double *rb = rbuf + (int64_t)(kh + j) * rows;
        ubuf[r] = 0.70710678118654752440 * (rbuf[(int64_t)j * rows + r] + jcomp(rb, r));
      }
      CTA_SYNC();
      emit_mode(outb, acc, rows, ubuf);
      PAR_FOR(one, 1) jb.e_out[acc] = 0.5;
      CTA_SYNC();
      ++acc;
    }
  }
  PAR_FOR(idx, k4 * rows) {
    const int c = idx / rows, r = idx - c * rows;
    jb.V[(int64_t)c * jb.ld + r] = outb[(int64_t)c * rows + r];
  }
  PAR_FOR(one, 1) {
    *jb.status = (int)red[4];
    *jb.kh_out = kh;
  }
}

// ---------------------------------------------------------------------------------------------
// K11: batched complex Pfaffians
// ---------------------------------------------------------------------------------------------
constexpr int PF_WARPS = 8;
constexpr int PF_PER_WARP = 16;            // entries per warp and CTA pass
constexpr int PF_CTA_ENTRIES = PF_WARPS * PF_PER_WARP;

TMF_DEVICE int pf_ctz64(uint64_t x) {
#if defined(TMF_HOSTSIM)
  return __builtin_ctzll(x);
#else
  return __ffsll((long long)x) - 1;
#endif
}

struct PfMeta {
  double pr, pi;      // running Pfaffian
  double tr, ti;      // 1 / pivot
  int p, zero;
  int idx[32];
};

TMF_GLOBAL pfaffians_kernel(const tmf_pf_block *blocks, const int *cta_prefix, int nblocks, int mmax,
                            int nmax) {
  int lo = 0, hi = nblocks;
  const int cta = BLOCK_ID;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cta_prefix[mid] <= cta) lo = mid; else hi = mid;
  }
  const tmf_pf_block blk = blocks[lo];
  const int64_t total = (int64_t)blk.n_bra * blk.n_ket;
  const int64_t e0 = (int64_t)(cta - cta_prefix[lo]) * PF_CTA_ENTRIES;
  const int m = blk.n1 + blk.n2, nn = blk.m;
  const int ldt = mmax + 1;                 // odd-ish stride against bank conflicts

  DYN_SMEM(unsigned char, raw);
  double *Nre = reinterpret_cast<double *>(raw);            // nmax * nmax
  double *Nim = Nre + (size_t)nmax * nmax;
  double *tiles = Nim + (size_t)nmax * nmax;                // PF_WARPS * 2 * mmax * ldt
  PfMeta *metas = reinterpret_cast<PfMeta *>(tiles + (size_t)PF_WARPS * 2 * mmax * ldt);

  PAR_FOR(idx, nn * nn) {
    Nre[idx] = blk.N[2 * idx];
    Nim[idx] = blk.N[2 * idx + 1];
  }
  CTA_SYNC();

  PF_WARP_FOR(w, PF_WARPS) {
    double *Are = tiles + (size_t)w * 2 * mmax * ldt;
    double *Aim = Are + (size_t)mmax * ldt;
    PfMeta *mt = metas + w;
    for (int q = 0; q < PF_PER_WARP; ++q) {
      const int64_t ent = e0 + (int64_t)q * PF_WARPS + w;
      if (ent >= total) break;
      const int a = (int)(ent / blk.n_ket), c = (int)(ent - (int64_t)a * blk.n_ket);
      // idx = ket excitations (low bits) ++ bra excitations (high bits), ascending (pfaffian.py:1468-1474)
      PF_LANE_FOR(l) if (l == 0) {
        uint64_t mask = blk.ket_masks[c] | blk.bra_masks[a];
        int t = 0;
        while (mask && t < 32) {
          mt->idx[t++] = pf_ctz64(mask);
          mask &= mask - 1;
        }
        mt->pr = blk.scale;
        mt->pi = 0.0;
        mt->zero = (m & 1);
      }
      PF_WSYNC();
      PF_LANE_FOR(l) {
        for (int e = l; e < m * m; e += 32) {
          const int r = e / m, s = e - r * m;
          const int src = mt->idx[r] * nn + mt->idx[s];
          Are[r * ldt + s] = (r == s) ? 0.0 : Nre[src];
          Aim[r * ldt + s] = (r == s) ? 0.0 : Nim[src];
        }
      }
      PF_WSYNC();
      for (int k = 0; k + 1 < m; k += 2) {
        // pivot: largest |A[k][j]|, j > k
        PF_LANE_FOR(l) if (l == 0) {
          double best = -1.0;
          int p = k + 1;
          for (int j = k + 1; j < m; ++j) {
            const double v = Are[k * ldt + j] * Are[k * ldt + j] + Aim[k * ldt + j] * Aim[k * ldt + j];
            if (v > best) { best = v; p = j; }
          }
          mt->p = p;
          const double xr = Are[k * ldt + p], xi = Aim[k * ldt + p];
          if (best <= 0.0) {
            mt->zero = 1;
            mt->tr = mt->ti = 0.0;
          } else {
            const double pr = mt->pr * xr - mt->pi * xi, pi = mt->pr * xi + mt->pi * xr;
            const double sg = (p != k + 1) ? -1.0 : 1.0;
            mt->pr = sg * pr;
            mt->pi = sg * pi;
            mt->tr = xr / best;       // 1 / (xr + i xi)
            mt->ti = -xi / best;
          }
        }
        PF_WSYNC();
        const int p = mt->p;
        if (p != k + 1) {   // symmetric interchange of rows / columns k+1 and p
          PF_LANE_FOR(l) {
            for (int s = l; s < m; s += 32) {
              double t = Are[(k + 1) * ldt + s]; Are[(k + 1) * ldt + s] = Are[p * ldt + s]; Are[p * ldt + s] = t;
              t = Aim[(k + 1) * ldt + s]; Aim[(k + 1) * ldt + s] = Aim[p * ldt + s]; Aim[p * ldt + s] = t;
            }
          }
          PF_WSYNC();
          PF_LANE_FOR(l) {
            for (int r = l; r < m; r += 32) {
              double t = Are[r * ldt + k + 1]; Are[r * ldt + k + 1] = Are[r * ldt + p]; Are[r * ldt + p] = t;
              t = Aim[r * ldt + k + 1]; Aim[r * ldt + k + 1] = Aim[r * ldt + p]; Aim[r * ldt + p] = t;
            }
          }
          PF_WSYNC();
        }
        // A[i][j] += tau_i A[j][k+1] - A[i][k+1] tau_j,  tau_i = A[k][i] / A[k][k+1],  i, j > k+1
        const int h = m - k - 2;
        if (h > 0) {
          PF_LANE_FOR(l) {
            const double tr = mt->tr, ti = mt->ti;
            for (int e = l; e < h * h; e += 32) {
              const int i = k + 2 + e / h, j = k + 2 + e % h;
              if (i == j) continue;
              const double air = Are[k * ldt + i], aii = Aim[k * ldt + i];
              const double ajr = Are[k * ldt + j], aji = Aim[k * ldt + j];
              const double tir = air * tr - aii * ti, tii = air * ti + aii * tr;   // tau_i
              const double tjr = ajr * tr - aji * ti, tji = ajr * ti + aji * tr;   // tau_j
              const double cjr = Are[j * ldt + k + 1], cji = Aim[j * ldt + k + 1]; // A[j][k+1]
              const double cir = Are[i * ldt + k + 1], cii = Aim[i * ldt + k + 1]; // A[i][k+1]
              Are[i * ldt + j] += (tir * cjr - tii * cji) - (cir * tjr - cii * tji);
              Aim[i * ldt + j] += (tir * cji + tii * cjr) - (cir * tji + cii * tjr);
            }
          }
          PF_WSYNC();
        }
      }
      PF_LANE_FOR(l) if (l == 0) {
        const bool z = mt->zero != 0;
        blk.out[2 * ent] = z ? 0.0 : mt->pr;
        blk.out[2 * ent + 1] = z ? 0.0 : mt->pi;
      }
      PF_WSYNC();
    }
  }
}

static size_t pf_smem_bytes(int mmax, int nmax) {
  return sizeof(double) * (2 * (size_t)nmax * nmax + (size_t)PF_WARPS * 2 * mmax * (mmax + 1)) +
         sizeof(PfMeta) * PF_WARPS + 64;
}

}  // namespace tmf

extern "C" int64_t tmf_pair_tmp_doubles(int rows) {
  return (int64_t)rows * (tmf::PAIR_MAXC + 1 + tmf::PAIR_MAXC / 2 + TMF_MAX_MODES);
}

extern "C" int tmf_pfaffian_pair_modes(const tmf_pair_job *jobs_host, int njobs, double half_tol,
                                       void *desc_dev, void *stream) {
  using namespace tmf;
  if (njobs <= 0) return TMF_OK;
  for (int j = 0; j < njobs; ++j)
    if (jobs_host[j].k4 > TMF_MAX_MODES || (jobs_host[j].k4 & 3)) {
      set_error("tmf_pfaffian_pair_modes: 4k must be a multiple of 4 and <= TMF_MAX_MODES");
      return TMF_ERR_VALUE;
    }
  int rc = copy_h2d(desc_dev, jobs_host, sizeof(tmf_pair_job) * (size_t)njobs, stream);
  if (rc) return rc;
  const size_t smem = sizeof(double) * (PAIR_MAXC + 2 * PAIR_MAXC * 33 + 2 * PAIR_MAXC + 8);
  return launch_t("pair_modes", pair_modes_kernel, njobs, 256, smem, stream,
                  reinterpret_cast<const tmf_pair_job *>(desc_dev), 1e-9, 1e-13, half_tol);
}

extern "C" int64_t tmf_pf_desc_bytes(int nblocks) {
  return tmf::align256(64 * (int64_t)nblocks) + tmf::align256(4 * (int64_t)(nblocks + 1)) + 256;
}

extern "C" int tmf_pfaffians_blocks(const tmf_pf_block *blocks_host, int nblocks, void *desc_dev,
                                    void *stream) {
  using namespace tmf;
  if (nblocks <= 0) return TMF_OK;
  std::vector<int> prefix(nblocks + 1, 0);
  int mmax = 1, nmax = 1;
  for (int b = 0; b < nblocks; ++b) {
    const tmf_pf_block &k = blocks_host[b];
    if (k.m > 64 || k.n1 + k.n2 > 32 || k.n1 < 0 || k.n2 < 0) {
      set_error("tmf_pfaffians_blocks: contraction matrix > 64 or Pfaffian size > 32 not supported");
      return TMF_ERR_VALUE;
    }
    const int64_t total = (int64_t)k.n_bra * k.n_ket;
    prefix[b + 1] = prefix[b] + (int)((total + PF_CTA_ENTRIES - 1) / PF_CTA_ENTRIES);
    mmax = std::max(mmax, k.n1 + k.n2);
    nmax = std::max(nmax, k.m);
  }
  if (prefix[nblocks] == 0) return TMF_OK;
  unsigned char *d = static_cast<unsigned char *>(desc_dev);
  const size_t o_pref = align256(sizeof(tmf_pf_block) * (size_t)nblocks);
  int rc = copy_h2d(d, blocks_host, sizeof(tmf_pf_block) * (size_t)nblocks, stream);
  if (rc) return rc;
  rc = copy_h2d(d + o_pref, prefix.data(), sizeof(int) * (size_t)(nblocks + 1), stream);
  if (rc) return rc;
  return launch_t("pfaffians", pfaffians_kernel, prefix[nblocks], 32 * PF_WARPS, pf_smem_bytes(mmax, nmax), stream,
                  reinterpret_cast<const tmf_pf_block *>(d), reinterpret_cast<const int *>(d + o_pref), nblocks,
                  mmax, nmax);
}
