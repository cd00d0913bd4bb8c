// Host-side (CPU) integer/ordering logic of the Slater -> MPS path: best-first subset enumeration,
// Schmidt-vector tables and per-site planning.  These are the parts of the reference that are
// inherently sequential integer bookkeeping (schmidt_utils.py:211-324, slater.py:633-700,
// :760-825, :1027-1058, :1106-1141); everything that is floating-point heavy runs in the CUDA
// kernels.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/temfpy_b200.h"

namespace tmf {

struct TruncPar {
  int chi_max = -1;  // <0: unlimited
  double svd_min = 1e-6;
  double degeneracy_tol = 1e-12;
  std::vector<int> sectors;  // empty + !filter -> all sectors
  bool filter = false;
  // symmetrise mode weights that are equal within the eigenvalue accuracy before the enumeration
  // (hostlogic.cpp: snap_degenerate).  On by default for inputs with exactly degenerate (spin-pure)
  // modes; off = the reference's literal behaviour (the cut inside such a multiplet is rounding noise).
  bool snap = true;
  bool is_sector(int q) const;
};

// schmidt_utils.py:211-324.  Returns sums (heap order, truncated) and masks.
void lowest_sums(const double *a, int k, double base, const TruncPar &tp, int filled_left,
                 int filled_right, std::vector<double> &sums, std::vector<uint64_t> &sets,
                 int *n_checked);

struct BondVectors {
  int k = 0, filled_left = 0;
  std::vector<uint64_t> masks;  // sorted stably by left charge
  std::vector<double> lam;      // un-normalised Schmidt values
  std::vector<int> charge;      // left charge of every vector
  std::vector<int> sec_q, sec_start;  // sector table (sec_start has one extra entry)
};

// symmetrises weights that are equal within the accuracy of the eigenvalues (see hostlogic.cpp)
void snap_degenerate(double *a, const double *e, int k);

// slater.py:633-700 on top of lowest_sums; e = left eigenvalues of the entangled modes.
void bond_vectors(const double *e, int k, int filled_left, const TruncPar &tp, BondVectors &out);

struct SitePlan {
  tmf_site_plan h{};
  std::vector<int> bra_cols, ket_cols;
  std::vector<double> bra_sign, ket_sign;
  std::vector<uint64_t> bra_masks, ket_masks;
  std::vector<int> row_p, row_alpha;
  std::vector<int> blocks;  // 6 ints per block
};

// Throws std::runtime_error on inconsistent input.
void site_plan(int mode, int n_bra, int n_ket, int k_bra, int f_bra, int nferm_bra, int chi_bra,
               const uint64_t *masks_bra, const int *charge_bra, int k_ket, int f_ket,
               int nferm_ket, int chi_ket, const uint64_t *masks_ket, const int *charge_ket,
               SitePlan &out);

// Device-resident result of the enumeration kernel (enumerate.cu): tables stay in the caller's workspace.
struct EnumResident {
  bool resident = false;
  int nb = 0, hw = 0;
  int64_t cap = 0;
  uint64_t *masks_dev = nullptr;     // nb x cap
  double *lam_dev = nullptr;         // nb x cap
  int *charge_dev = nullptr;         // nb x cap
  int *head_dev = nullptr;           // nb x hw: chi, sectors, status, pops, sec_q[66], sec_start[67]
  std::vector<int> head_host;        // host copy of the heads
};

void set_error(const std::string &msg);

// Runs f(0) .. f(n-1) on a persistent pool of host threads (at most max_threads of them plus the caller,
// dynamic scheduling); rethrows the first exception.  Spawning fresh threads for every stage of every
// pipeline chunk cost more than the stages themselves.
void pool_for(int n, int max_threads, const std::function<void(int)> &f);

}  // namespace tmf
