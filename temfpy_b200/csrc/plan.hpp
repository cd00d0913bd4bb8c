// Device-side site planning (plan.cu): descriptor and entry points shared with the chain driver.
#pragma once
#include <cstddef>
#include <cstdint>

namespace tmf {

constexpr int PLAN_MAX_SECTORS = 66;   // = ENUM_MAX_SECTORS (charges of a bond span at most k + 1 <= 65 values)
constexpr int PLAN_MAX_BLOCKS = 66;
constexpr int PLAN_MAX_ORB = 66;       // physical + 64 entangled modes + edge vector
constexpr int PLAN_HDR_INTS = 24;      // tmf_site_plan (18 ints), [18] status, padding

struct PlanJob {
  const uint64_t *masks_b, *masks_k;   // occupation masks of the two bonds (enumeration output, device)
  const int *charge_b, *charge_k;      // (unused by the kernel today; kept for symmetry with the host planner)
  const int *head_b, *head_k;          // enumeration heads: [0] chi, [1] sectors, [4..] sec_q, [4 + 66..] sec_start
  uint64_t *bra_masks, *ket_masks;     // out: occupation of the sometimes rows / cols (capacity 2 cap / cap)
  int *cols;                           // out: bra_cols[PLAN_MAX_ORB] | ket_cols[PLAN_MAX_ORB]
  double *signs;                       // out: bra_sign[PLAN_MAX_ORB] | ket_sign[PLAN_MAX_ORB]
  int *hdr;                            // out: PLAN_HDR_INTS
  int *blocks;                         // out: 6 ints per block, capacity PLAN_MAX_BLOCKS
  int mode, k_bra, k_ket, df, n_bra, n_ket, f_bra, f_ket;
};

size_t plan_smem_bytes();
int plan_sites_device(const PlanJob *jobs_dev, int nsites, void *stream);

}  // namespace tmf
