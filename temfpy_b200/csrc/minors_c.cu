// K10 for complex Slater determinants: all minors det(S[rows(alpha)][:, cols(beta)]) of every charge block with a
// complex128 sometimes matrix (slater.py:828-869 with complex `sometimes_matrix`; the c128 variant of
// tmf_minors_blocks).  Same shared-elimination algorithm as minors.cu -- the bra row is row-reduced once
// (Gauss-Jordan, pivoting along the row), every ket costs a determinant of the size of its column exchange --
// in its plain form: one warp per bra row, reduced matrix in a shared-memory tile, pivoted LU for the small
// determinants.  Complex inputs are the rarer case; the register-resident / binned fast path stays real-only.
#include "cplx.hpp"

namespace tmf {

constexpr int MC_ROWS = 16;     // bra rows per CTA
constexpr int MC_WARPS = 4;
constexpr int MC_DMAX = 16;     // simultaneous column exchanges handled (as minors.cu)

#if defined(TMF_HOSTSIM)
#define MC_LANE_FOR(l) for (int l = 0; l < 32; ++l)
#define MC_WARP_FOR(w, W) for (int w = 0; w < (W); ++w)
#define MC_WSYNC() ((void)0)
TMF_DEVICE int mc_popc(uint64_t x) { return __builtin_popcountll(x); }
TMF_DEVICE int mc_ctz(uint64_t x) { return __builtin_ctzll(x); }
#else
#define MC_LANE_FOR(l) for (int l = (threadIdx.x & 31), l##_once = 1; l##_once; l##_once = 0)
#define MC_WARP_FOR(w, W) for (int w = (threadIdx.x >> 5), w##_once = 1; w##_once && w < (W); w##_once = 0)
#define MC_WSYNC() __syncwarp()
TMF_DEVICE int mc_popc(uint64_t x) { return __popcll(x); }
TMF_DEVICE int mc_ctz(uint64_t x) { return __ffsll((long long)x) - 1; }
#endif

TMF_DEVICE uint64_t mc_prefix_parity(uint64_t x) {
  uint64_t p = x << 1;
  p ^= p << 1; p ^= p << 2; p ^= p << 4; p ^= p << 8; p ^= p << 16; p ^= p << 32;
  return p;
}

struct MCMeta {
  uint64_t c0, pp0;
  cplx scale, inv;
  int pc, pad_;
  int colrow[64];
  double cand[64];
};

// determinant of the d x d matrix buf (row-major, ld = MC_DMAX) by LU with partial pivoting
TMF_DEVICE cplx mc_det(cplx *m, int d) {
  cplx det = cmake(1.0);
  for (int j = 0; j < d; ++j) {
    int p = j;
    double best = cabs2(m[j * MC_DMAX + j]);
    for (int i = j + 1; i < d; ++i) {
      const double a = cabs2(m[i * MC_DMAX + j]);
      if (a > best) { best = a; p = i; }
    }
    if (best == 0.0) return cmake(0.0);
    if (p != j) {
      det = cneg(det);
      for (int c = 0; c < d; ++c) { const cplx t = m[j * MC_DMAX + c]; m[j * MC_DMAX + c] = m[p * MC_DMAX + c]; m[p * MC_DMAX + c] = t; }
    }
    const cplx piv = m[j * MC_DMAX + j];
    det = cmul(det, piv);
    const cplx ip = cinv(piv);
    for (int i = j + 1; i < d; ++i) {
      const cplx l = cmul(m[i * MC_DMAX + j], ip);
      for (int c = j + 1; c < d; ++c) m[i * MC_DMAX + c] = csub(m[i * MC_DMAX + c], cmul(l, m[j * MC_DMAX + c]));
    }
  }
  return det;
}

// one tensor entry: signed determinant of Y[C0 \ C, C \ C0]
TMF_DEVICE cplx mc_entry(const cplx *x, int smax, const int *colrow, uint64_t c0, uint64_t pp0, uint64_t cm, uint64_t ppk) {
  uint64_t mu = c0 & ~cm, de = cm & ~c0;
  const int d = mc_popc(de);
  if (d == 0) return cmake(1.0);
  if (d > MC_DMAX) return cmake(NAN, NAN);
  const int par = (mc_popc(mu & pp0) ^ mc_popc(de & ppk)) & 1;
  int rr[MC_DMAX], cc[MC_DMAX];
  for (int i = 0; i < d; ++i) {
    const int m = mc_ctz(mu), e = mc_ctz(de);
    mu &= mu - 1;
    de &= de - 1;
    rr[i] = colrow[m] * smax;
    cc[i] = e;
  }
  cplx val;
  if (d == 1) {
    val = x[rr[0] + cc[0]];
  } else if (d == 2) {
    val = csub(cmul(x[rr[0] + cc[0]], x[rr[1] + cc[1]]), cmul(x[rr[0] + cc[1]], x[rr[1] + cc[0]]));
  } else {
    cplx buf[MC_DMAX * MC_DMAX];
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) buf[i * MC_DMAX + j] = x[rr[i] + cc[j]];
    val = mc_det(buf, d);
  }
  return par ? cneg(val) : val;
}

TMF_GLOBAL minors_c_kernel(const tmf_minor_block *blocks, const int *cta_prefix, int nblocks, int nmax, int smax,
                           int nkmax) {
  int lo = 0, hi = nblocks;
  const int cta = BLOCK_ID;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cta_prefix[mid] <= cta) lo = mid; else hi = mid;
  }
  const tmf_minor_block blk = blocks[lo];
  const int row0 = (cta - cta_prefix[lo]) * MC_ROWS;
  const int nrows = (blk.n_bra - row0 < MC_ROWS) ? (blk.n_bra - row0) : MC_ROWS;
  const int n = blk.minor, sk = blk.s_ket, sb = blk.s_bra;
  DYN_SMEM(unsigned char, raw);
  const int ldS = sb | 1;
  cplx *Ssm = reinterpret_cast<cplx *>(raw);                     // smax * (smax | 1)
  cplx *Xall = Ssm + (size_t)smax * (smax | 1);                  // MC_WARPS * nmax * smax
  MCMeta *metas = reinterpret_cast<MCMeta *>(Xall + (size_t)MC_WARPS * nmax * smax);
  uint64_t *kmask = reinterpret_cast<uint64_t *>(metas + MC_WARPS);   // nkmax
  uint64_t *kpp = kmask + nkmax;                                 // nkmax
  const cplx *Sg = reinterpret_cast<const cplx *>(blk.S);
  const cplx det_always = blk.det ? cmake(blk.det[0], blk.det[1]) : cmake(1.0);
  cplx *out = reinterpret_cast<cplx *>(blk.out);
  PAR_FOR(idx, sb * sk) {
    const int c = idx / sb, r = idx - c * sb;
    Ssm[c * ldS + r] = Sg[idx];
  }
  PAR_FOR(c, blk.n_ket) {
    const uint64_t km = blk.ket_masks[c];
    kmask[c] = km;
    kpp[c] = mc_prefix_parity(km);
  }
  CTA_SYNC();
  if (n == 0) {
    PAR_FOR(idx, nrows * blk.n_ket) {
      const int a = idx / blk.n_ket, c = idx - a * blk.n_ket;
      out[(int64_t)(row0 + a) * blk.n_ket + c] = det_always;
    }
    return;
  }
  MC_WARP_FOR(w, MC_WARPS) {
    cplx *x = Xall + (size_t)w * nmax * smax;
    MCMeta *mt = metas + w;
    for (int a = w; a < nrows; a += MC_WARPS) {
      const uint64_t rmask = blk.bra_masks[row0 + a];
      MC_LANE_FOR(l) {
        for (int c = l; c < sk; c += 32) {
          uint64_t rm = rmask;
          int r = 0;
          while (rm) {
            const int bit = mc_ctz(rm);
            rm &= rm - 1;
            x[r * smax + c] = Ssm[(size_t)c * ldS + bit];
            ++r;
          }
        }
        if (l == 0) { mt->c0 = 0; mt->scale = det_always; }
      }
      MC_WSYNC();
      for (int t = 0; t < n; ++t) {
        MC_LANE_FOR(l) {
          for (int c = l; c < sk; c += 32) mt->cand[c] = ((mt->c0 >> c) & 1) ? -1.0 : cabs2(x[t * smax + c]);
        }
        MC_WSYNC();
        MC_LANE_FOR(l) if (l == 0) {
          double best = -1.0;
          int bc = 0;
          for (int c = 0; c < sk; ++c)
            if (mt->cand[c] > best) { best = mt->cand[c]; bc = c; }
          const cplx pv = x[t * smax + bc];
          mt->pc = bc;
          mt->inv = cinv(pv);
          mt->scale = cmul(mt->scale, pv);
          mt->c0 |= (1ull << bc);
          mt->colrow[bc] = t;
        }
        MC_WSYNC();
        MC_LANE_FOR(l) {
          const int pc = mt->pc;
          const cplx inv = mt->inv;
          for (int c = l; c < sk; c += 32) {
            if (c == pc) continue;
            const cplx u = cmul(x[t * smax + c], inv);
            x[t * smax + c] = u;
            for (int r = 0; r < n; ++r)
              if (r != t) x[r * smax + c] = csub(x[r * smax + c], cmul(x[r * smax + pc], u));
          }
        }
        MC_WSYNC();
      }
      MC_LANE_FOR(l) if (l == 0) {
        uint64_t c0 = mt->c0, seen = 0;
        int inv = 0;
        while (c0) {
          const int c = mc_ctz(c0);
          c0 &= c0 - 1;
          const int r = mt->colrow[c];
          inv += mc_popc(seen >> (r + 1));
          seen |= (1ull << r);
        }
        if (inv & 1) mt->scale = cneg(mt->scale);
        mt->pp0 = mc_prefix_parity(mt->c0);
      }
      MC_WSYNC();
      cplx *orow = out + (int64_t)(row0 + a) * blk.n_ket;
      MC_LANE_FOR(l) {
        const uint64_t c0 = mt->c0, pp0 = mt->pp0;
        for (int c = l; c < blk.n_ket; c += 32)
          orow[c] = cmul(mt->scale, mc_entry(x, smax, mt->colrow, c0, pp0, kmask[c], kpp[c]));
      }
      MC_WSYNC();
    }
  }
}

static size_t minors_c_smem_bytes(int nmax, int smax, int nkmax) {
  return sizeof(cplx) * ((size_t)smax * (smax | 1) + (size_t)MC_WARPS * nmax * smax) + sizeof(MCMeta) * MC_WARPS +
         16 * (size_t)nkmax + 64;
}

}  // namespace tmf

// c128 variant of tmf_minors_blocks: S, det and out are complex (re, im interleaved); block descriptors as there.
extern "C" int tmf_minors_blocks_c(const tmf_minor_block *blocks_host, int nblocks, void *desc_dev, void *stream) {
  using namespace tmf;
  if (nblocks <= 0) return TMF_OK;
  std::vector<int> prefix(nblocks + 1, 0);
  int nmax = 1, smax = 1, nkmax = 1;
  for (int b = 0; b < nblocks; ++b) {
    const tmf_minor_block &k = blocks_host[b];
    if (k.s_bra > 64 || k.s_ket > 64 || k.minor > 32 || k.minor > k.s_ket || k.minor > k.s_bra) {
      set_error("tmf_minors_blocks_c: sometimes matrix > 64 or minor size > 32 not supported");
      return TMF_ERR_VALUE;
    }
    const int ctas = (k.n_bra > 0 && k.n_ket > 0) ? (k.n_bra + MC_ROWS - 1) / MC_ROWS : 0;
    prefix[b + 1] = prefix[b] + ctas;
    nmax = std::max(nmax, k.minor);
    smax = std::max(smax, std::max(k.s_bra, k.s_ket));
    nkmax = std::max(nkmax, k.n_ket);
  }
  nkmax = (nkmax + 3) & ~3;
  if (prefix[nblocks] == 0) return TMF_OK;
  const size_t smem = minors_c_smem_bytes(nmax, smax, nkmax);
  if (smem > 220 * 1024) {
    set_error("tmf_minors_blocks_c: block too large for the shared-memory tiles");
    return TMF_ERR_VALUE;
  }
  unsigned char *d = static_cast<unsigned char *>(desc_dev);
  const size_t o_pref = align256(sizeof(tmf_minor_block) * (size_t)nblocks);
  int rc = copy_h2d(d, blocks_host, sizeof(tmf_minor_block) * (size_t)nblocks, stream);
  if (rc) return rc;
  rc = copy_h2d(d + o_pref, prefix.data(), sizeof(int) * (size_t)(nblocks + 1), stream);
  if (rc) return rc;
  return launch_t("minors_c", minors_c_kernel, prefix[nblocks], 32 * MC_WARPS, smem, stream,
                  reinterpret_cast<const tmf_minor_block *>(d), reinterpret_cast<const int *>(d + o_pref), nblocks, nmax,
                  smax, nkmax);
}
