// a7 / a8 / a10 planning on the device: the integer bookkeeping of one site from the two bond tables the
// enumeration kernel left in device memory -- no copy of the tables to the host, no host threads.
//
// reference: slater._select_orbitals (slater.py:760-825: always / sometimes / never classes, reordering
// signs), MPSTensorData.from_schmidt_vectors (:1027-1058: physical orbital, stable row sort by pipe charge),
// to_npc_array (:1106-1141: charge blocks).  Same results, bit for bit, as tmf::site_plan (hostlogic.cpp),
// which stays the specification and the path of the legacy / host-enumerated chains; the CPU suite compares
// the two on the kernel simulator.
//
// One CTA per site.  Orbitals in the nested layout (siteprep.cu): bra = physical + entangled modes, ket =
// entangled modes + edge vector (stored-V column k_ket) when the filled count grows by one.
#include "cta.hpp"
#include "plan.hpp"

namespace tmf {

static_assert(sizeof(PlanJob) == 128, "plan descriptor must be 128 bytes");

namespace {
struct POrb {
  short kind, idx, cls, col;   // kind: 0 entangled, 1 filled (edge vector), 2 physical
};
TMF_DEVICE int plan_popc(uint64_t x) {
#if defined(TMF_HOSTSIM)
  return __builtin_popcountll(x);
#else
  return __popcll(x);
#endif
}
TMF_DEVICE int plan_ctz(uint64_t x) {
#if defined(TMF_HOSTSIM)
  return __builtin_ctzll(x);
#else
  return __ffsll((long long)x) - 1;
#endif
}
// occupation mask over the `rest` orbitals of one Schmidt vector (MaskMap of hostlogic.cpp)
TMF_DEVICE uint64_t plan_map(uint64_t m, int p, bool complement, uint64_t kmask, uint64_t constant, uint64_t phys,
                             const uint64_t *bit_of) {
  if (complement) m = ~m;
  m &= kmask;
  uint64_t r = constant | (p == 1 ? phys : 0ull);
  while (m) {
    r |= bit_of[plan_ctz(m)];
    m &= m - 1;
  }
  return r;
}
}  // namespace

TMF_GLOBAL site_plan_kernel(const PlanJob *jobs) {
  const PlanJob jb = jobs[BLOCK_ID];
  const int mode = jb.mode, kb_modes = jb.k_bra, kk_modes = jb.k_ket;
  const int chi_b = jb.head_b[0], chi_k = jb.head_k[0];
  const int ns_b = jb.head_b[1], ns_k = jb.head_k[1];
  const int *secq_b = jb.head_b + 4, *secs_b = jb.head_b + 4 + PLAN_MAX_SECTORS;
  const int *secq_k = jb.head_k + 4, *secs_k = jb.head_k + 4 + PLAN_MAX_SECTORS;
  DYN_SMEM(unsigned char, raw);
  uint64_t *red = reinterpret_cast<uint64_t *>(raw);            // 4 * 64: all_b, any_b, all_k, any_k partials
  uint64_t *bit_b = red + 4 * 64;                               // 64
  uint64_t *bit_k = bit_b + 64;                                 // 64
  uint64_t *misc64 = bit_k + 64;                                // 8: const_b, phys_b, const_k, phys_k, all_b, any_b, all_k, any_k
  int *cls_q = reinterpret_cast<int *>(misc64 + 8);             // 2 * PLAN_MAX_SECTORS + 2: class charges
  int *cls_base = cls_q + 2 * PLAN_MAX_SECTORS + 2;             // class start rows
  int *cls_c0 = cls_base + 2 * PLAN_MAX_SECTORS + 2;            // p = 0 rows of the class
  int *cls_c1 = cls_c0 + 2 * PLAN_MAX_SECTORS + 2;              // p = 1 rows of the class
  int *sec_cls0 = cls_c1 + 2 * PLAN_MAX_SECTORS + 2;            // PLAN_MAX_SECTORS: class of the p = 0 rows of bra sector j
  int *sec_cls1 = sec_cls0 + PLAN_MAX_SECTORS;                  // class of its p = 1 rows
  int *misc = sec_cls1 + PLAN_MAX_SECTORS;                      // 8: n_classes, n_blocks, status
  int *blk = misc + 8;                                          // 6 * PLAN_MAX_BLOCKS (staged, then copied out)

  // ---- all / any over the occupation masks of the two bonds ---------------------------------------
  PAR_FOR(t, 64) {
    uint64_t a = ~0ull, y = 0ull;
    for (int i = t; i < chi_b; i += 64) { const uint64_t m = jb.masks_b[i]; a &= m; y |= m; }
    red[t] = a; red[64 + t] = y;
    a = ~0ull; y = 0ull;
    for (int i = t; i < chi_k; i += 64) { const uint64_t m = jb.masks_k[i]; a &= m; y |= m; }
    red[128 + t] = a; red[192 + t] = y;
  }
  CTA_SYNC();
  PAR_FOR(w, 4) {
    uint64_t v = red[w * 64];
    for (int t = 1; t < 64; ++t) v = (w & 1) ? (v | red[w * 64 + t]) : (v & red[w * 64 + t]);
    misc64[4 + w] = v;
  }
  CTA_SYNC();

  // ---- orbital classes, selection, signs, mask maps, row classes (one thread) ------------------------
  PAR_FOR(one, 1) {
    int status = 0;
    POrb ob[TMF_MAX_MODES + 2], ok[TMF_MAX_MODES + 2];
    int nob = 0, nok = 0;
    auto ent_cls = [&](uint64_t all, uint64_t any, int i) -> short {
      const bool a = (all >> i) & 1, y = (any >> i) & 1;      // occupation on the LEFT
      if (mode == 0) return a ? 2 : (y ? 1 : 0);              // left vectors: occupied iff bit set
      return !y ? 2 : (a ? 0 : 1);                            // right vectors: occupied iff bit clear
    };
    const uint64_t all_b = misc64[4], any_b = misc64[5], all_k = misc64[6], any_k = misc64[7];
    // orbitals in the reference's column order (slater.py:355-368, :1030-1051)
    if (mode == 0) {
      for (int i = 0; i < kb_modes; ++i) ob[nob++] = POrb{0, (short)i, ent_cls(all_b, any_b, i), (short)i};
      ob[nob++] = POrb{2, 0, 1, -1};
      for (int t = 0; t < jb.df; ++t) ok[nok++] = POrb{1, (short)t, 2, (short)(kk_modes + t)};
      for (int i = 0; i < kk_modes; ++i) ok[nok++] = POrb{0, (short)i, ent_cls(all_k, any_k, i), (short)i};
    } else {
      ob[nob++] = POrb{2, 0, 1, -1};
      for (int i = kb_modes - 1; i >= 0; --i) ob[nob++] = POrb{0, (short)i, ent_cls(all_b, any_b, i), (short)i};
      for (int i = kk_modes - 1; i >= 0; --i) ok[nok++] = POrb{0, (short)i, ent_cls(all_k, any_k, i), (short)i};
      for (int t = 0; t < jb.df; ++t) ok[nok++] = POrb{1, (short)t, 2, (short)(kk_modes + t)};
    }
    // select (slater.py:792-821): always / sometimes lists with the reordering signs
    int kab = 0, kak = 0;
    for (int i = 0; i < nob; ++i) kab += ob[i].cls == 2;
    for (int i = 0; i < nok; ++i) kak += ok[i].cls == 2;
    const int k = kab < kak ? kab : kak;                       // slater.py:1069
    auto emit = [&](const POrb *orbs, int n, int kside, int *cols, double *signs, uint64_t *bit_of, uint64_t &constant,
                    uint64_t &phys, int &n_rest) {
      // output order: [all always | sometimes]; `rest` (rows / cols of the sometimes matrix) in the reference
      // order: left  = surplus always (after the first k) + sometimes, right = sometimes + surplus always (before the last k)
      int ia = 0, is = 0, n_before = 0;
      int n_some = 0;
      for (int i = 0; i < n; ++i) n_some += orbs[i].cls == 1;
      const int surplus = kside - k;
      n_rest = surplus + n_some;
      for (int i = 0; i < 64; ++i) bit_of[i] = 0ull;
      constant = phys = 0ull;
      auto rest_bit = [&](const POrb &o, int t) {
        if (t >= 64) return;                                   // reported as status 2 below
        if (o.kind == 2) phys |= (1ull << t);
        else if (o.kind == 1) constant |= (1ull << t);
        else bit_of[o.idx] |= (1ull << t);
      };
      for (int i = 0; i < n; ++i) {
        const POrb &o = orbs[i];
        if (o.cls == 2) {
          cols[ia] = o.col; signs[ia] = 1.0;
          if (mode == 0) { if (ia >= k) rest_bit(o, ia - k); }
          else { if (ia < surplus) rest_bit(o, n_some + ia); }
          ++ia; ++n_before;
        } else if (o.cls == 1) {
          const int expo = (mode == 0) ? (kside - n_before) : n_before;
          cols[kside + is] = o.col; signs[kside + is] = (expo & 1) ? -1.0 : 1.0;
          rest_bit(o, (mode == 0) ? surplus + is : is);
          ++is;
        }
      }
    };
    int s_bra = 0, s_ket = 0;
    uint64_t cb = 0, pb = 0, ck = 0, pk = 0;
    emit(ob, nob, kab, jb.cols, jb.signs, bit_b, cb, pb, s_bra);
    emit(ok, nok, kak, jb.cols + PLAN_MAX_ORB, jb.signs + PLAN_MAX_ORB, bit_k, ck, pk, s_ket);
    misc64[0] = cb; misc64[1] = pb; misc64[2] = ck; misc64[3] = pk;
    if (s_bra > 64 || s_ket > 64) status = 2;
    // ---- row classes: pipe charge Q = q + dq * p, p = 0 rows before p = 1 rows inside a class (stable sort) ----
    const int dq = (mode == 0) ? 1 : -1;
    int nc = 0;
    {
      int i0 = 0, i1 = 0;      // merge of the ascending lists q_j and q_j + dq
      while (i0 < ns_b || i1 < ns_b) {
        const int q0 = (i0 < ns_b) ? secq_b[i0] : (1 << 30), q1 = (i1 < ns_b) ? secq_b[i1] + dq : (1 << 30);
        const int Q = q0 < q1 ? q0 : q1;
        cls_q[nc] = Q; cls_c0[nc] = 0; cls_c1[nc] = 0;
        if (q0 == Q) { cls_c0[nc] = secs_b[i0 + 1] - secs_b[i0]; sec_cls0[i0] = nc; ++i0; }
        if (q1 == Q) { cls_c1[nc] = secs_b[i1 + 1] - secs_b[i1]; sec_cls1[i1] = nc; ++i1; }
        ++nc;
      }
      int base = 0;
      for (int c = 0; c < nc; ++c) { cls_base[c] = base; base += cls_c0[c] + cls_c1[c]; }
    }
    misc[0] = nc;
    misc[2] = status;
    // header
    int *h = jb.hdr;
    h[0] = mode; h[1] = 1; h[2] = jb.n_bra; h[3] = jb.n_ket; h[4] = kb_modes; h[5] = kk_modes;
    h[6] = jb.f_bra; h[7] = jb.f_ket; h[8] = k; h[9] = s_bra; h[10] = s_ket; h[11] = 2 * chi_b;
    h[12] = chi_b; h[13] = chi_k; h[14] = 0; h[15] = 0; h[16] = kab; h[17] = kak;   // qtotal = 0: both bonds of one chain
  }
  CTA_SYNC();
  if (misc[2] != 0) {
    PAR_FOR(one, 1) { jb.hdr[18] = misc[2]; jb.hdr[14] = 0; }
    return;
  }
  // ---- occupation masks of the bra rows (in sorted row order) and of the ket vectors ----------------
  {
    const bool complement = (mode != 0);
    const uint64_t kmask_b = (kb_modes >= 64) ? ~0ull : ((1ull << kb_modes) - 1ull);
    const uint64_t kmask_k = (kk_modes >= 64) ? ~0ull : ((1ull << kk_modes) - 1ull);
    const uint64_t cb = misc64[0], pb = misc64[1], ck = misc64[2], pk = misc64[3];
    PAR_FOR(a, chi_b) {
      int j = 0;
      while (j + 1 < ns_b && secs_b[j + 1] <= a) ++j;
      const int within = a - secs_b[j];
      const uint64_t m = jb.masks_b[a];
      const int c0 = sec_cls0[j], c1 = sec_cls1[j];
      jb.bra_masks[cls_base[c0] + within] = plan_map(m, 0, complement, kmask_b, cb, pb, bit_b);
      jb.bra_masks[cls_base[c1] + cls_c0[c1] + within] = plan_map(m, 1, complement, kmask_b, cb, pb, bit_b);
    }
    PAR_FOR(c, chi_k) jb.ket_masks[c] = plan_map(jb.masks_k[c], 0, complement, kmask_k, ck, pk, bit_k);
  }
  CTA_SYNC();
  // ---- charge blocks (slater.py:1132-1141) ---------------------------------------------------------
  PAR_FOR(one, 1) {
    const int nc = misc[0];
    int nb = 0, c = 0;
    for (int j = 0; j < ns_k; ++j) {
      const int q_ket = secq_k[j], c0 = secs_k[j], c1 = secs_k[j + 1];
      const int Q = q_ket;                                   // + qtotal * qc, qtotal = 0 (slater.py:1092, :1134)
      while (c < nc && cls_q[c] < Q) ++c;
      if (c < nc && cls_q[c] == Q && cls_c0[c] + cls_c1[c] > 0 && c1 > c0) {
        int *b = blk + 6 * nb;
        b[0] = cls_base[c]; b[1] = cls_c0[c] + cls_c1[c]; b[2] = c0; b[3] = c1 - c0;
        b[4] = plan_popc(jb.ket_masks[c0]); b[5] = q_ket;
        ++nb;
      }
    }
    misc[1] = nb;
  }
  CTA_SYNC();
  const int nb = misc[1];
  // equal particle numbers inside every block (slater.py:847-855)
  for (int b = 0; b < nb; ++b) {
    const int *bl = blk + 6 * b;
    PAR_FOR(i, bl[1] + bl[3]) {
      const uint64_t m = (i < bl[1]) ? jb.bra_masks[bl[0] + i] : jb.ket_masks[bl[2] + i - bl[1]];
      if (plan_popc(m) != bl[4]) misc[2] = 1;
    }
  }
  PAR_FOR(i, 6 * nb) jb.blocks[i] = blk[i];
  CTA_SYNC();
  PAR_FOR(one, 1) { jb.hdr[14] = nb; jb.hdr[18] = misc[2]; }
}

size_t plan_smem_bytes() {
  return sizeof(uint64_t) * (4 * 64 + 64 + 64 + 8) +
         sizeof(int) * (4 * (2 * PLAN_MAX_SECTORS + 2) + 2 * PLAN_MAX_SECTORS + 8 + 6 * PLAN_MAX_BLOCKS) + 64;
}

int plan_sites_device(const PlanJob *jobs_dev, int nsites, void *stream) {
  if (nsites <= 0) return TMF_OK;
  return launch_t("site_plan", site_plan_kernel, nsites, 256, plan_smem_bytes(), stream, jobs_dev);
}

}  // namespace tmf
