// One-sided Jacobi (small SVD / symmetric eigenproblem of one CTA) and the ranking helper, shared by the mode
// extraction (modes_kernels.cuh) and the canonical-form sweep (canon.cu).  PAR_FOR / CTA_SYNC style: the CPU
// simulator executes the same source.
#pragma once
#include "cta.hpp"
#include <cstdio>

namespace tmf {

constexpr int JAC_MAX_SWEEPS = 40;

// ---------------------------------------------------------------------------------------------
// one-sided Jacobi on the columns of G (n x n, ldg) with accumulation in J (n x n, ldj).
// On exit the columns of G are mutually orthogonal, G_out = G_in * J, J orthogonal.
//   * SVD use:  G = R  -> singular values = column norms, right singular vectors = J
//   * eigen use: G = symmetric PSD matrix -> eigenvalues = column norms, eigenvectors = J
// Parallel round-robin ordering; `rot` (n/2 * 2 doubles), `part` (n/2 * 33 * 3) and `flag` are
// shared scratch.  n may be odd (a bye is inserted).
// ---------------------------------------------------------------------------------------------
TMF_DEVICE int jacobi_onesided(double *G, int ldg, double *J, int ldj, int n, double *rot,
                               double *part, int *flag) {
  if (n < 2) return 0;
  int sweeps_done = 0;
  const int np = (n + 1) & ~1;  // padded to even; index np-1 == n is a bye when n is odd
  const int half = np / 2;
  for (int sweep = 0; sweep < JAC_MAX_SWEEPS; ++sweep) {
    PAR_FOR(one, 1) *flag = 0;
    CTA_SYNC();
    for (int round = 0; round < np - 1; ++round) {
#if !defined(TMF_HOSTSIM)
      // CUDA path: one group of GW lanes per column pair (GW = 8 for n <= 64: four pairs share a warp).  The
      // rotation scalars are computed redundantly by every lane, ~45 FP64 instructions per warp and round
      // whatever the number of pairs in the warp, and with 6 warps per scheduler the Jacobi kernels are bound
      // by the FP64 issue rate of the SM -- four pairs per warp cost a quarter of it per pair.  Pairs of a
      // round touch disjoint columns, so the dot products (shuffle reductions inside the group), the rotation
      // and the column updates of a pair need no CTA barrier; one __syncthreads per round separates the pairings.
      {
        const int GW = (n <= 64) ? 8 : (n <= 128 ? 16 : 32);
        const int gpw = 32 / GW;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
        const int sub = lane / GW, gl = lane - sub * GW;
        for (int base = warp * gpw; base < half; base += nwarp * gpw) {
          const int pr = base + sub;
          int p = round + pr, q = round + np - 1 - pr;      // circle method without integer division
          if (p >= np - 1) p -= np - 1;
          if (q >= np - 1) q -= np - 1;
          if (pr == 0) p = np - 1;
          const bool active = pr < half && p < n && q < n;   // (bye of an odd n, tail of the last warp)
          double *gp = G + (int64_t)(active ? p : 0) * ldg, *gq = G + (int64_t)(active ? q : 0) * ldg;
          double a = 0.0, b = 0.0, c = 0.0;
          if (active)
            for (int r = gl; r < n; r += GW) {
              const double x = gp[r], y = gq[r];
              a += x * x;
              b += y * y;
              c += x * y;
            }
          for (int o = GW >> 1; o > 0; o >>= 1) {             // uniform for the whole warp
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
          }
          // a column with norm < 1e-15 (all our matrices have norm O(1)) is numerically null: rounding
          // noise of the rotations with the other columns keeps changing it by O(1) of its own size, so
          // pairs involving it would never meet the relative criterion -- skip them (as LAPACK's dgesvj)
          if (active && c * c > 1e-30 * a * b && a > 1e-30 && b > 1e-30) {
            // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (b - a) / (2 c), written without the
            // two IEEE divisions and the IEEE square root (~25 FP64 instructions each when emulated).  rsqrt and a
            // Newton-refined reciprocal are accurate to a few ulp, ample for a rotation angle (cs^2 + sn^2 = 1
            // holds to the same few ulp; J is renormalised at the end).
            const double d = b - a, h = 2.0 * c;
            const double q2 = d * d + h * h;
            const double rs = rsqrt(q2);
            const double den = fabs(d) + q2 * rs;            // |d| + sqrt(d^2 + h^2)
            double rc;
            if (den > 1e-30 && den < 1e30) {
              rc = (double)__frcp_rn((float)den);
              rc = rc * (2.0 - den * rc);
              rc = rc * (2.0 - den * rc);
              rc = rc * (2.0 - den * rc);
            } else {
              rc = 1.0 / den;
            }
            const double t = ((d >= 0.0) == (h >= 0.0) ? fabs(h) : -fabs(h)) * rc;
            const double cs = rsqrt(1.0 + t * t), sn = cs * t;
            if (gl == 0) *flag = 1;
            for (int r = gl; r < n; r += GW) {
              const double x = gp[r], y = gq[r];
              gp[r] = cs * x - sn * y;
              gq[r] = sn * x + cs * y;
            }
            if (J != nullptr) {
              double *jp = J + (int64_t)p * ldj, *jq = J + (int64_t)q * ldj;
              for (int r = gl; r < n; r += GW) {
                const double x = jp[r], y = jq[r];
                jp[r] = cs * x - sn * y;
                jq[r] = sn * x + cs * y;
              }
            }
          }
        }
        __syncthreads();
        continue;
      }
#endif
      // simulator path (same arithmetic, shared-memory staging instead of shuffles)
      // circle method: position i of the top row meets position i of the bottom row
      PAR_FOR(item, half * 32) {
        int pr = item >> 5, lane = item & 31;
        int p = (pr == 0) ? np - 1 : (round + pr) % (np - 1);
        int q = (round + np - 1 - pr) % (np - 1);
        double a = 0.0, b = 0.0, c = 0.0;
        if (p < n && q < n) {
          const double *gp = G + (int64_t)p * ldg, *gq = G + (int64_t)q * ldg;
          for (int r = lane; r < n; r += 32) {
            double x = gp[r], y = gq[r];
            a += x * x;
            b += y * y;
            c += x * y;
          }
        }
        double *pp = part + (pr * 33 + lane) * 3;
        pp[0] = a; pp[1] = b; pp[2] = c;
      }
      CTA_SYNC();
      PAR_FOR(pr, half) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int l = 0; l < 32; ++l) {
          const double *pp = part + (pr * 33 + l) * 3;
          a += pp[0]; b += pp[1]; c += pp[2];
        }
        double cs = 1.0, sn = 0.0;
        if (c * c > 1e-30 * a * b && a > 1e-30 && b > 1e-30) {
          double zeta = (b - a) / (2.0 * c);
          double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          cs = 1.0 / sqrt(1.0 + t * t);
          sn = cs * t;
          *flag = 1;  // benign race: every writer stores 1
        }
        rot[2 * pr] = cs;
        rot[2 * pr + 1] = sn;
      }
      CTA_SYNC();
      PAR_FOR(item, half * n) {
        int pr = item / n, r = item - pr * n;
        int p = (pr == 0) ? np - 1 : (round + pr) % (np - 1);
        int q = (round + np - 1 - pr) % (np - 1);
        double cs = rot[2 * pr], sn = rot[2 * pr + 1];
        if (p < n && q < n && sn != 0.0) {
          double *gp = G + (int64_t)p * ldg + r, *gq = G + (int64_t)q * ldg + r;
          double x = *gp, y = *gq;
          *gp = cs * x - sn * y;
          *gq = sn * x + cs * y;
          if (J != nullptr) {
            double *jp = J + (int64_t)p * ldj + r, *jq = J + (int64_t)q * ldj + r;
            x = *jp; y = *jq;
            *jp = cs * x - sn * y;
            *jq = sn * x + cs * y;
          }
        }
      }
      CTA_SYNC();
    }
    const int any = *flag;
    sweeps_done = sweep + 1;
    CTA_SYNC();
#if defined(TMF_HOSTSIM) && defined(TMF_DEBUG_SWEEPS)
    if (!any || sweep == JAC_MAX_SWEEPS - 1) fprintf(stderr, "jacobi n=%d sweeps=%d\n", n, sweep + 1);
#endif
    if (!any) break;
  }
  // the product of ~n * sweeps plane rotations drifts from orthonormality by ~1e-14: renormalise the
  // columns of J (first-order repair; the residual non-orthogonality only enters at second order)
  if (J != nullptr) {
    PAR_FOR(c, n) {
      double s = 0.0;
      for (int r = 0; r < n; ++r) s += J[(int64_t)c * ldj + r] * J[(int64_t)c * ldj + r];
      rot[c] = (s > 0.0) ? 1.0 / sqrt(s) : 1.0;
    }
    CTA_SYNC();
    PAR_FOR(idx, n * n) {
      int c = idx / n, r = idx - c * n;
      J[(int64_t)c * ldj + r] *= rot[c];
      G[(int64_t)c * ldg + r] *= rot[c];
    }
    CTA_SYNC();
  }
  return sweeps_done;
}
inline size_t jacobi_scratch_doubles(int n) {
  int half = ((n + 1) & ~1) / 2;
  return (size_t)half * 2 + (size_t)half * 33 * 3 + 8;
}

// ranks `key[0..n)` by decreasing value (ties by index); flagged entries only.  rank_out[i] = -1
// for unflagged entries.  Returns nothing; count of flagged entries in *count.
TMF_DEVICE void rank_desc(const double *key, const int *flagged, int n, int *rank_out, int *count) {
  PAR_FOR(i, n) {
    int r = -1;
    if (flagged[i]) {
      r = 0;
      for (int j = 0; j < n; ++j)
        if (flagged[j] && (key[j] > key[i] || (key[j] == key[i] && j < i))) ++r;
    }
    rank_out[i] = r;
  }
  PAR_FOR(one, 1) {
    int c = 0;
    for (int j = 0; j < n; ++j) c += flagged[j] ? 1 : 0;
    *count = c;
  }
  CTA_SYNC();
}

}  // namespace tmf
