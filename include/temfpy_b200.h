/*
 * temfpy_b200 -- C ABI of the B200-native mean-field -> MPS hot path.
 *
 * The reference (temfpy/temfpy, pure Python) has no FFI boundary of its own: its boundary for this
 * path is the public Python API (slater.py / schmidt_utils.py / ...).  Every entry point below
 * replaces the arithmetic behind one reference function; the citation after "replaces:" is the
 * reference file:line (relative to /root/reference/src/temfpy/).  The Python shim in
 * temfpy_b200/ binds these symbols with ctypes (see INTEGRATION.md for the stub a reference
 * maintainer would add).
 *
 * Conventions
 *  - every matrix is column-major with an explicit leading dimension; the correlation matrix is
 *    Hermitian so its NumPy (row-major) buffer can be passed unchanged (real symmetric case);
 *  - "dev" pointers are device pointers into caller-owned buffers (PyTorch allocations), "host"
 *    pointers are plain host memory; nothing here allocates device memory;
 *  - `stream` is a cudaStream_t passed as void*; all device work is enqueued on it and the call
 *    returns without synchronising unless stated;
 *  - return value: 0 = ok, <0 = error class (TMF_ERR_*), message via tmf_last_error().
 */
#ifndef TEMFPY_B200_H
#define TEMFPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMF_OK 0
#define TMF_ERR_VALUE (-1)     /* -> ValueError   (bad argument, capacity exceeded)          */
#define TMF_ERR_ASSERT (-2)    /* -> AssertionError (reference asserts, e.g. slater.py:394)  */
#define TMF_ERR_RUNTIME (-3)   /* -> RuntimeError (CUDA failure, no device)                  */

#define TMF_SIDE_L 0
#define TMF_SIDE_R 1
#define TMF_MAX_MODES 64       /* entangled modes per bond handled by the 64-bit occupation masks */

/* library --------------------------------------------------------------------------------- */
int tmf_version(void);
const char *tmf_last_error(void);
/* 1 if this build runs kernels on a CUDA device, 0 for the host simulator used by CPU tests. */
int tmf_is_cuda(void);
int tmf_device_count(void);

/* K2 -- correlation matrix build.  replaces: slater.py:1177 (C = v @ HT(v)).
 * C (L x L, ldc) = Phi * Phi^T with Phi row-major L x N (row stride ldphi, NumPy layout); real
 * symmetric, FP64 DMMA tiles. */
int tmf_corr_build(const double *phi_dev, int L, int N, int ldphi, double *C_dev, int ldc,
                   void *stream);

/* General grouped FP64 GEMM used by the mode extraction (exposed for tests / profiling).
 * Jobs are described on the host; descriptors are copied into `desc_dev` (njobs * 128 bytes). */
typedef struct tmf_gemm_job {
  const double *A, *B;
  double *C;
  const int *a_idx, *b_idx;            /* optional column gathers for op(A) rows / op(B) cols   */
  const double *row_scale, *col_scale; /* optional epilogue scalings                            */
  int M, N, K;
  int lda, ldb, ldc;
  int transA, transB;                  /* op(A) is M x K, op(B) is K x N                        */
  int a_row_off, b_row_off;            /* row offset applied along K inside A / B               */
  double alpha, beta;
  int pad_[4];
} tmf_gemm_job;
int64_t tmf_gemm_desc_bytes(int njobs);    /* size of desc_dev for the call below */
int tmf_gemm_grouped(const tmf_gemm_job *jobs_host, int njobs, void *desc_dev, void *stream);

/* Input check of slater.C_to_MPS (the reference's centre-bond assertion eL + eR = 1, slater.py:404, only holds for a
 * projector): max |(C C - C)[i, j]| over the first `rows` rows of the symmetric C (pitch ldc).  work_dev: rows * L + 1
 * doubles, the result is work_dev[rows * L]; desc_dev: tmf_gemm_desc_bytes(1). */
int tmf_projector_defect(const double *C_dev, int L, int ldc, int rows, double *work_dev, void *desc_dev,
                         void *stream);

/* K3/K4 -- per-bond Schmidt-mode extraction.  replaces: slater.py:324-375 (diag_and_separate:
 * eigh :347, split :350, reorder :353-370) for every (bond, side) job of a chain at once.
 *
 * For job j the diagonal block A = C[:x,:x] (side L) or C[x:,x:] (side R) of the projector C is
 * decomposed into k_j entangled eigenpairs (cutoff < e < 1-cutoff) and an orthonormal basis of
 * the "filled" eigenspace (e >= 1-cutoff).  Output per job, column-major with ld = n_j at
 * V_dev + v_off[j]:  columns [0,k) entangled modes ordered by decreasing *left* eigenvalue
 * (the order of the reference's `e` array), columns [k, k+f) filled basis.
 *   e_dev[j*TMF_MAX_MODES + i]  left-eigenvalue of entangled mode i
 *   info_dev[4*j + {0,1,2,3}] = k, f, status (0 ok, 1 sketch rank exhausted -> rerun wider), n
 * `work_dev` must hold tmf_slater_modes_workspace() bytes.  r_sketch in {64,128}. */
int64_t tmf_slater_modes_workspace(int L, int njobs, const int *job_x, const int *job_side,
                                   int r_sketch);
int tmf_slater_modes_batched(const double *C_dev, int L, int ldc, int njobs, const int *job_x,
                             const int *job_side, double cutoff, int r_sketch,
                             const int64_t *v_off, double *V_dev, double *e_dev, int *info_dev,
                             void *work_dev, int64_t work_bytes, void *stream);

/* Nested form used by the chain driver: no basis of the filled space.  Column k of a job's V slot
 * receives the *edge vector* g = P_F e_edge / |P_F e_edge| (P_F = projector onto the filled space of the
 * block, e_edge = unit vector of the block's site next to the cut) -- the one direction by which the filled
 * space grows when the block grows by a site; info[1] = f = round(tr A - sum of the entangled
 * eigenvalues); edge_dev[2 j] = |P_F e_edge|^2, edge_dev[2 j + 1] = rounding remainder of f.
 * V slots need tmf_slater_modes_slot_cols() columns (ld = n).  replaces: slater.py:347-370 as above;
 * the filled eigenvectors the reference keeps (:355-368) are never needed, see tmf_site_nested_batched. */
int64_t tmf_slater_modes_slot_cols(int L, int x, int side, int r_sketch, int nested);
int tmf_slater_modes_nested(const double *C_dev, int L, int ldc, int njobs, const int *job_x,
                            const int *job_side, double cutoff, int r_sketch, const int64_t *v_off,
                            double *V_dev, double *e_dev, int *info_dev, double *edge_dev,
                            void *work_dev, int64_t work_bytes, void *stream);

/* Complex Slater determinants (slater.py:1150-1180 keeps a complex C; :347 then diagonalises complex Hermitian
 * blocks).  The library takes the re/im-interleaved real embedding Cemb (2L x 2L real symmetric projector,
 * Cemb[2i+a, 2j+b] = [[Re, -Im], [Im, Re]] of C_ij) with the cuts at 2x: the real kernels extract the modes (every
 * complex eigenvector is a doubly degenerate real pair), a pairing kernel picks the k complex modes -- stored
 * as interleaved complex columns in the same slots -- and the embedded edge vector is the complex one.
 * On return info[4j] = k complex modes, info[4j+1] = f, e[0..k) the left eigenvalues. */
int tmf_slater_modes_nested_emb(const double *Cemb_dev, int L2, int ldc, int njobs, const int *job_x2,
                                const int *job_side, double cutoff, int r_sketch, const int64_t *v_off,
                                double *V_dev, double *e_dev, int *info_dev, double *edge_dev,
                                void *work_dev, int64_t work_bytes, void *stream);
/* complex form of tmf_slater_pair_bond (utils.py:19-96 + slater.py:410 with complex mode matrices) */
int64_t tmf_slater_pair_bond_c_workspace(int L, int k);
int tmf_slater_pair_bond_c(const double *Cemb_dev, int ldc, int L, int x, int k, const double *e_host,
                           double degeneracy_tol, double *VL_dev, double *VR_dev, void *work_dev,
                           int64_t work_bytes, void *stream);

/* K4 -- pairing of the left and right entangled modes of a bond whose two sides were both extracted.
 * replaces: utils.py:19-96 (block_svd) as called from slater.py:407, and the odd-index sign flips of
 * slater.py:410.  VL (x rows) / VR (L-x rows): stored mode matrices of the two jobs; their first k
 * columns are rotated in place (GEMMs V_L^T C_LR V_R on the device, SVD of the degenerate k x k
 * groups on the host -> one synchronisation).  e_host: the k left eigenvalues (host pointer).
 * Called by tmf_chain_tensors for the centre bond and by the iMPS driver for its two cuts. */
int64_t tmf_slater_pair_bond_workspace(int L, int k);
int tmf_slater_pair_bond(const double *C_dev, int ldc, int L, int x, int k, const double *e_host,
                         double degeneracy_tol, double *VL_dev, double *VR_dev, void *work_dev,
                         int64_t work_bytes, void *stream);

/* K6/K7 -- best-first enumeration of the most probable occupation subsets (HOST, multi-threaded
 * over bonds).  replaces: schmidt_utils.py:211-324 (lowest_sums), :99-185 (StoppingCondition
 * __call__/truncate), slater.py:673-689 (sort by charge, sector table, Schmidt values).
 *
 * Single-problem form (mirrors lowest_sums): a[k] and base = sum(a[a<0]) are supplied by the
 * caller (NumPy computes them in the reference).  sectors: list of allowed charges or NULL.
 * Outputs in heap order: sums_out[n], sets_out[n] (bit i = entry i of `a` selected). */
int tmf_lowest_sums(const double *a, int k, double base, int chi_max, double svd_min,
                    double degeneracy_tol, const int *sectors, int n_sectors, int filled_left,
                    int filled_right, int cap, double *sums_out, uint64_t *sets_out, int *n_out,
                    int *n_checked);

/* Batched bond form: for bond b, e[b*TMF_MAX_MODES + i], k[b], filled_left[b].
 * Outputs per bond (capacity `cap` rows each): masks sorted stably by left charge, un-normalised
 * Schmidt values, left charges, chi[b], and the sector table (sec_q, sec_start; sec_n[b] sectors,
 * capacity TMF_MAX_MODES+1 per bond). */
int tmf_bond_vectors_batched(int nbonds, const double *e, const int *k, const int *filled_left,
                             int chi_max, double svd_min, double degeneracy_tol,
                             const int *sectors, int n_sectors, int cap, uint64_t *masks,
                             double *lam, int *charge, int *chi, int *sec_q, int *sec_start,
                             int *sec_n, int n_threads);

/* a7/a8 planning (HOST).  replaces the integer bookkeeping of slater.py:760-825
 * (_select_orbitals), :1027-1058 (physical orbital, stable row sort) and :1106-1141 (charge
 * blocks) for one site.  See temfpy_b200/csrc/plan.cpp for the exact field semantics. */
typedef struct tmf_site_plan {
  int mode;          /* 0 = left, 1 = right                                              */
  int physical;      /* 1: bra carries the site's orbital                                */
  int n_bra, n_ket;  /* number of sites the bra / ket orbitals live on                   */
  int k_bra, k_ket;  /* entangled modes of the two bonds                                 */
  int f_bra, f_ket;  /* filled orbitals of the two bonds                                 */
  int k_always;      /* size of the square "always" block (min of the two sides)         */
  int s_bra, s_ket;  /* rows / cols of the sometimes matrix                              */
  int n_rows;        /* bra rows (2*chi_bra with physical leg)                           */
  int chi_bra, chi_ket;
  int n_blocks;
  int qtotal;
  int ka_bra, ka_ket; /* always-occupied orbitals of the bra / ket side (k_always = min)    */
} tmf_site_plan;

/* Plans one site.  Inputs: the two bonds' masks (sorted, from tmf_bond_vectors_batched), charges.
 * Outputs (caller-allocated, capacities in brackets):
 *   bra_cols[k_bra+f_bra+1], ket_cols[k_ket+f_ket]: stored-V column index of every O row / col
 *       in the order [all always orbitals (ka_bra / ka_ket) | sometimes orbitals], -1 = physical;
 *   bra_sign / ket_sign: the reordering signs;
 *   bra_masks[n_rows], ket_masks[chi_ket]: occupation of the sometimes rows / cols (bit t = row t);
 *   row_p[n_rows], row_alpha[n_rows]: physical index and bond index of every bra row;
 *   blocks[n_blocks*6]: bra_row_start, n_bra_rows, ket_start, n_ket, minor size, charge q_ket. */
int tmf_slater_site_plan(int mode, int n_bra, int n_ket, int k_bra, int f_bra, int nferm_bra,
                         int chi_bra, const uint64_t *masks_bra, const int *charge_bra, int k_ket,
                         int f_ket, int nferm_ket, int chi_ket, const uint64_t *masks_ket,
                         const int *charge_ket, tmf_site_plan *plan, int *bra_cols,
                         double *bra_sign, int *ket_cols, double *ket_sign, uint64_t *bra_masks,
                         uint64_t *ket_masks, int *row_p, int *row_alpha, int *blocks);

/* K8+K9 -- overlap of the two mode bases and Schur complement, batched over sites.
 * replaces: slater.py:1071 (O = HT(v_bra) @ v_ket) and :1073-1090 (det_always, sometimes
 * matrix).  Site descriptors (host) are copied to desc_dev.  For site s:
 *   O (rows = bra_cols order, cols = ket_cols order) is formed in work; k = min(ka_bra, ka_ket)
 *   elimination steps with pivoting over the always orbitals of the larger side; S receives the
 *   (s_bra x s_ket) Schur complement in the reference's row / column order (surplus always
 *   orbitals after the sometimes ones for right tensors, before them for left tensors),
 *   det = +-det(always block) (the sign is a global phase of the site tensor). */
typedef struct tmf_site_job {
  const double *Vb, *Vk;       /* stored mode matrices of the bra / ket bond-side             */
  const int *bra_cols, *ket_cols;       /* device arrays (rows / cols of O), -1 = physical     */
  const double *bra_sign, *ket_sign;    /* device arrays                                      */
  double *O;                   /* (ka_bra + sb) x (ka_ket + sk) workspace                     */
  double *S;                   /* output s_bra x s_ket, ld = s_bra                            */
  double *det;                 /* output scalar                                               */
  int ldb, ldk;
  int n_bra, n_ket;            /* sites                                                       */
  int mode, physical;
  int ka_bra, ka_ket;          /* always orbitals of each side                                */
  int sb, sk;                  /* sometimes orbitals of each side (sb includes the physical)  */
  /* pad_[0] (internal): 1 = O and *det were prepared by the nested-projector kernel;
   * pad_[1]: 1 = report a vanishing pivot of the always block (|pivot| < 1e-9) as *det = NaN     */
  int emb;                     /* 1: Pfaffian path -- rows are re/im-interleaved Majorana components
                                  (4 per site); bra_cols -1..-4 = emb(w), J emb(w) of the physical
                                  site's lower mode (c^+ row, pfaffian.py:1667-1688) and the same
                                  for its upper mode (c row)                                  */
  int pad_[3];
} tmf_site_job;
int64_t tmf_site_desc_bytes(int nsites);   /* size of desc_dev for the call below */
int tmf_site_overlap_schur_batched(const tmf_site_job *jobs_host, int nsites, void *desc_dev,
                                   void *stream);

/* K8 + K9, nested-projector form (chain driver).  replaces: slater.py:1071-1090 for two neighbouring
 * bonds of the *same* correlation matrix.  The filled spaces of the two blocks are nested up to the
 * truncation threshold (F_ket = F_bra (+) edge vector, or F_ket = F_bra), so their elimination has the
 * closed form  S = Y_b^T Y_k - (Z_b^T Z_k)(1 - Z_k^T Z_k)^-1  with Y = explicit orbitals (physical +
 * entangled modes + edge vector) and Z = their leaks into the other side's filled space, all obtained
 * from one pass over the n x k mode matrices (see siteprep.cu); the always-occupied *entangled* orbitals
 * are then eliminated as in tmf_site_overlap_schur_batched.  Site jobs as above with V slots in the nested
 * layout ([entangled | edge vector]; column codes of the plan: i < k entangled, k = edge vector). */
typedef struct tmf_nested_job {
  const double *e_bra, *e_ket; /* device: left eigenvalues of the bra / ket entangled modes           */
  const double *a_col;         /* device: C[site, bra block] (n_bra contiguous doubles)               */
  const double *c_edge;        /* device: &C[site, site]                                              */
  int k_bra, k_ket;            /* entangled modes of the two bonds                                    */
  int df;                      /* f_ket - f_bra: 1 if the ket's filled space has the edge vector      */
  int pad_[5];
} tmf_nested_job;
int tmf_site_nested_batched(const tmf_site_job *jobs_host, const tmf_nested_job *njobs_host,
                            int nsites, void *desc_dev, void *stream);

/* complex128 form (complex Slater determinants): V slots hold interleaved complex columns (ldb / ldk in complex
 * elements), a_col / c_edge point into the embedded matrix (row 2e holds conj(C[e, :]) = C[:, e]^T), O / S / det
 * are complex; the always-occupied entangled orbitals are eliminated in the same kernel (<= 32 modes per bond). */
int tmf_site_nested_c_batched(const tmf_site_job *jobs_host, const tmf_nested_job *njobs_host,
                              int nsites, void *desc_dev, void *stream);

/* K10 -- all minors of all charge blocks of all sites.  replaces: slater.py:828-869
 * (_tensor_block: gather + batched det) and the det_always scaling of :1137.
 * Block descriptors (host) are copied to desc_dev.  For block b:
 *   out[a * n_ket + c] = det_always * det(S[rows(bra_masks[a])][:, cols(ket_masks[c])])
 * with S (s_bra x s_ket, column-major ld = s_bra).  Rows/cols of a minor are the set bits in
 * ascending order (slater.py:857-867). */
typedef struct tmf_minor_block {
  const double *S;
  const double *det;           /* device scalar (det_always) or NULL for 1.0                  */
  const uint64_t *bra_masks;   /* n_bra masks                                                 */
  const uint64_t *ket_masks;   /* n_ket masks                                                 */
  double *out;                 /* n_bra x n_ket row-major                                     */
  int s_bra, s_ket, n_bra, n_ket, minor, pad_;
} tmf_minor_block;
int64_t tmf_minor_desc_bytes(int nblocks); /* size of desc_dev for the call below */
int tmf_minors_blocks(const tmf_minor_block *blocks_host, int nblocks, void *desc_dev,
                      void *stream);
/* c128 variant (slater.py:857-869 with a complex sometimes matrix): S, det and out are complex, re/im interleaved */
int tmf_minors_blocks_c(const tmf_minor_block *blocks_host, int nblocks, void *desc_dev, void *stream);


/* ---- Pfaffian (Bogoliubov) path --------------------------------------------------------------
 * The Majorana-basis Nambu correlation matrix C_M = 1/2 + iA (pfaffian.py:269-273) is handed to the
 * library as its real representation P' (re/im interleaved, 4L x 4L real symmetric projector).  The
 * per-bond eigenproblems (pfaffian.py:789) then run through tmf_slater_modes_batched on P' with the
 * cuts at 4x; the overlap / Schur stage through tmf_site_overlap_schur_batched with `emb` = 1. */

/* K4p -- complex modes out of the real eigenvectors of P'.  replaces: the complex eigenvectors that
 * eigh returns at pfaffian.py:789 and the Nambu completion of :886/:891.
 * For job j (one CTA): V (rows x >=k4, ld) holds the k4 = 4k raw entangled columns of
 * tmf_slater_modes_batched (ordered by decreasing left eigenvalue, e_raw[k4]); on return columns
 * 4a..4a+3 are emb(w_a), J emb(w_a), emb(conj w_a), J emb(conj w_a) for the k modes with block
 * eigenvalue e_out[a] <= 1/2 (ascending).  Eigenvalues within half_tol of 1/2 (pfaffian.py:802-816,
 * half_tol = degeneracy_tol) span the complexification of a real null space: a real orthonormal
 * basis r_1..r_2kh is extracted and paired as w_j = (r_j + i r_{kh+j}) / sqrt 2 (pfaffian.py:884);
 * which real vectors are paired is a gauge choice (the reference shuffles them with a fixed random
 * orthogonal matrix, :867-874).  tmp: tmf_pair_tmp_doubles(rows) doubles of workspace.
 * status: 0 ok, 2 bad k4, 3 odd multiplicity, 4 missing plane, 5 asymmetric 1/2 spectrum,
 * 6 1/2 eigenvectors cannot be made real. */
typedef struct tmf_pair_job {
  double *V;
  double *tmp;
  const double *e_raw;
  double *e_out;
  int *status;
  int *kh_out;                 /* number of modes with eigenvalue 1/2 (the last kh of the k modes) */
  int rows, ld, k4, side;
} tmf_pair_job;
int64_t tmf_pair_tmp_doubles(int rows);
int tmf_pfaffian_pair_modes(const tmf_pair_job *jobs_host, int njobs, double half_tol,
                            void *desc_dev /* 64*njobs */, void *stream);

/* K11 -- all Pfaffians of a (bra excitation number, ket excitation number) block.
 * replaces: pfaffian.py:1429-1479 (_tensor_block: gather + one pfapack call per entry, :1425).
 *   out[a * n_ket + c] = scale * Pf(N[idx, idx]),  idx = bits(ket_masks[c]) ++ bits(bra_masks[a])
 * N: m x m complex128 (re, im interleaved), row-major, antisymmetric; bit t of a mask = index t of
 * N (ket modes occupy the low indices, pfaffian.py:1400-1408); n1 / n2 = excitations per bra / ket
 * mask; out is complex128 (interleaved), row-major n_bra x n_ket. */
typedef struct tmf_pf_block {
  const double *N;
  const uint64_t *bra_masks;
  const uint64_t *ket_masks;
  double *out;
  double scale;
  int m, n_bra, n_ket, n1, n2, pad_;
} tmf_pf_block;
int64_t tmf_pf_desc_bytes(int nblocks);
int tmf_pfaffians_blocks(const tmf_pf_block *blocks_host, int nblocks, void *desc_dev, void *stream);


/* K9p -- per-site finish of the Pfaffian site stage.  replaces: pfaffian.py:1339-1400 (singular values and inverse of
 * the U* block of V1^+ V2, blocks AA / BA / BB, the antisymmetric contraction matrix N) together with the sign / swap
 * fixes of :1665, :1708-1719, :915-916 and the centre-bond rotations of :855, applied to the Schur complement S that
 * tmf_site_overlap_schur_batched (emb = 1) left per site: column-major (sb + sur_b) x (sk + sur_k), surplus rows /
 * columns last for right tensors (mode 1).  One CTA per job.  out[0] = product, out[1] = smallest of the singular
 * values of U* (a vanishing one: vacua of opposite parity, :1355-1357); with want_n the m x m complex N (interleaved,
 * row-major, m = active ket + active bra modes, ket modes first in descending order, :1361-1374, :1400-1408) is
 * written to N.  idx1_mask / idx2_mask: bit t = bra / ket mode t is occupied in some Schmidt vector.  rot_up / rot_lo:
 * optional real 2 k2 x 2 k2 matrices (row-major) multiplied from the right onto the upper / lower ket pairs. */
typedef struct tmf_pf_site_job {
  const double *S;
  const double *rot_up, *rot_lo;
  double *N;
  double *out;
  double *work_;             /* unused (the kernel works in shared memory) */
  uint32_t idx1_mask, idx2_mask;
  int sb, sk, sur_b, sur_k, mode, k1, k2, fix, want_n, no_phys;  /* no_phys: no physical mode among the bra modes */
  double u_p, ket_sign;
  int pad_[4];
} tmf_pf_site_job;
int tmf_pfaffian_site_finish(const tmf_pf_site_job *jobs_host, int njobs, void *desc_dev /* 128 * njobs bytes */,
                             void *stream);

/* Chain driver -- replaces the per-site loop of slater.C_to_MPS (slater.py:1216-1353) for the
 * sites [site_lo, site_hi) of one chain (one call sequence per GPU; shards need no communication).
 *   create -> modes_sizes -> [caller allocates] -> modes (device + one D2H sync) -> enumerate (host)
 *   -> tensor_sizes -> [caller allocates] -> tensors (device, asynchronous) -> bond/site accessors.
 * Result layout: block b of site s is a dense row-major (n_bra_rows x n_ket) matrix at
 * out_dev + block_off[b]; its rows follow the reference's bra-pipe order (row_p, row_alpha). */
typedef struct tmf_chain tmf_chain;
tmf_chain *tmf_chain_create(int L, int ortho_center, int n_fermion, int chi_max, double svd_min,
                            double degeneracy_tol, const int *sectors, int n_sectors, int r_sketch,
                            int site_lo, int site_hi, int n_threads);
void tmf_chain_destroy(tmf_chain *c);
int tmf_chain_modes_sizes(tmf_chain *c, int64_t *q /* njobs, V doubles, workspace bytes */);
int tmf_chain_modes(tmf_chain *c, const double *C_dev, int ldc, double *V_dev, double *e_dev,
                    int *info_dev, void *work_dev, int64_t work_bytes, void *stream);
/* The same in two halves: enqueue all kernels (no host wait) / fetch the spectra (waits). */
int tmf_chain_modes_enqueue(tmf_chain *c, const double *C_dev, int ldc, double *V_dev, double *e_dev,
                            int *info_dev, void *work_dev, int64_t work_bytes, void *stream);
int tmf_chain_modes_finish(tmf_chain *c, const double *e_dev, const int *info_dev, void *stream);
int tmf_chain_enumerate(tmf_chain *c);
/* The same with the subset enumeration (schmidt_utils.py:211-324, slater.py:633-700) on the device:
 * one warp per bond; work_dev holds tmf_chain_enum_workspace(c) bytes.  Bit-identical tables. */
int64_t tmf_chain_enum_workspace(tmf_chain *c);
int tmf_chain_enumerate_dev(tmf_chain *c, void *work_dev, int64_t work_bytes, void *stream);
int tmf_chain_tensor_sizes(tmf_chain *c, int64_t *q /* plan bytes, O, S doubles, sites, blocks,
                                                       out doubles, max chi */);
int tmf_chain_tensors(tmf_chain *c, const double *C_dev, int ldc, double *V_dev, void *plan_dev,
                      int64_t plan_bytes, double *O_dev, double *S_dev, double *det_dev,
                      double *out_dev, void *stream);
int tmf_chain_bond(tmf_chain *c, int bond, int *q /* chi, k, filled_left, n_sectors, jobL, jobR,
                   fL, fR */, const double **lam, const int **charge, const uint64_t **masks,
                   const int **sec_q, const int **sec_start, const double **e);
int tmf_chain_site(tmf_chain *c, int site, tmf_site_plan *plan, const int **blocks,
                   const int64_t **block_off, const int **row_p, const int **row_alpha,
                   int64_t *offs /* O offset, S offset, det index */);
/* Bulk export of the host-side results of a shard (one call instead of one per bond / site).
 * Replaces the attribute reads of SchmidtVectors (slater.py:494-543) and MPSTensorData (:872-973) that
 * C_to_MPS performs for every bond and site (:1296-1346). */
int tmf_chain_bonds_sizes(tmf_chain *c, int64_t *q /* first bond, bonds, sum chi, sum sectors */);
int tmf_chain_bonds_export(tmf_chain *c, int64_t *chi_off, int *head /* k, filled_left, fL, fR */,
                           double *lam, int *charge, uint64_t *masks, int64_t *sec_off, int *sec_q,
                           int *sec_start, double *e);
int tmf_chain_sites_sizes(tmf_chain *c, int64_t *q /* sites, sum n_blocks, sum n_rows */);
int tmf_chain_sites_export(tmf_chain *c, tmf_site_plan *plans, int64_t *blk_off, int *blocks,
                           int64_t *block_off, int64_t *row_off, int *row_p, int *row_alpha);
/* (row_p / row_alpha may be NULL: the row order is a function of the bra bond's charges -- [p = 0 | p = 1]
 * stably sorted by charge +- p, slater.py:1053-1058 -- and callers can derive it lazily) */
int64_t tmf_chain_job_voff(tmf_chain *c, int job);
/* Options (call right after tmf_chain_create):
 *   TMF_OPT_SNAP   1 (default): mode weights equal within the eigenvalue accuracy are symmetrised
 *                  before the enumeration (spin-pure degenerate modes); 0: the reference's literal
 *                  behaviour (schmidt_utils.py:175-185 sees only degeneracies below degeneracy_tol).
 *   TMF_OPT_NESTED 1 (default): nested-projector site stage (no filled bases); 0: explicit filled
 *                  bases by pivoted Cholesky + overlap GEMM + blocked LU (also the automatic fallback). */
#define TMF_OPT_SNAP 1
#define TMF_OPT_NESTED 2
#define TMF_OPT_COMPLEX 4      /* 1: complex Slater determinant -- C_dev of modes / tensors is the 2L x 2L real embedding
                                * (pitch ldc doubles), r_sketch counts real columns (twice the complex modes), V /
                                * O / S / det / out buffers hold complex numbers (2 doubles per element) */
#define TMF_OPT_DEVICE_PLAN 3  /* 1 (default): site planning (slater.py:760-825, :1027-1058, :1106-1141) on the
                                * device from the resident enumeration tables (nested mode); 0: host threads */
#define TMF_OPT_PEER_OUT 5     /* 1: out_dev of tmf_chain_tensors is memory of another GPU (peer window, tmf_ipc_open):
                                * the minors kernel collects every tensor row in shared memory and writes it with
                                * whole 256-byte warp stores (NVLink carries each store as one packet) */
int tmf_chain_set_option(tmf_chain *c, int option, int value);
/* algorithmic flops of the reference's algorithm for this shard:
 * f[0] eigh, f[1] overlap GEMM, f[2] Schur, f[3] minors, f[4] number of minors */
int tmf_chain_flops(tmf_chain *c, double *f);

/* K14 -- Gutzwiller projection, one fused launch for all spin sites of an MPS.
 * replaces: the TeNPy arithmetic behind gutzwiller.py:227 (group_sites(2)) + :242 (iproject) + :244 (drop_charge)
 * (abrikosov) resp. :409, :424, :437-441 (abrikosov_ph).  A job is one surviving charge chain of one pair of
 * fermion sites:
 *     out[i * so_i + n] = row_scale[i] * sum_k A[i * sa_i + k * sa_k] * k_scale[k] * B[k * sb_k + n * sb_n] * col_scale[n]
 * (strides in elements; elements are f64, or interleaved c128 when `cplx_flag` is set; the three real scalings are
 * optional: the Schmidt values of the orthogonality centre).  A / B point into the block-sparse fermion tensors
 * where tmf_chain_tensors left them in HBM, `out` into the dense spin-site tensor T[vL, s, vR].  The call zero-fills
 * [out_dev, out_dev + out_bytes), uploads the descriptors to `desc_dev` (tmf_gutz_desc_bytes) and launches. */
typedef struct tmf_gutz_job {
  const void *A, *B;
  void *out;
  const double *k_scale, *row_scale, *col_scale;
  int64_t sa_i, sa_k, sb_k, sb_n, so_i;
  int m, k, n, pad_[7];
} tmf_gutz_job;
int64_t tmf_gutz_desc_bytes(const tmf_gutz_job *jobs_host, int njobs);
int tmf_gutzwiller_project(const tmf_gutz_job *jobs_host, int njobs, int cplx_flag, void *out_dev,
                           int64_t out_bytes, void *desc_dev, void *stream);

/* f3 -- canonical form of a finite charge-conserving MPS on the device.  replaces: TeNPy's
 * MPS.canonical_form_finite as called at gutzwiller.py:266 / :471 (left-to-right QR sweep, right-to-left SVD sweep,
 * block-wise in the charge sectors).  Tensors are dense T[vL, p, vR] (row-major doubles), the charges of a bond are
 * sorted (contiguous sectors), q(vL) + qp[p] = q(vR).  tmf_canon_create plans both sweeps from the sector tables
 * (dims0[L+1] bond dimensions, charges0 concatenated, qp[2]); tmf_canon_sizes: q[0] workspace bytes, q[1] doubles
 * of all output tensors, q[2] number of singular values (bonds 0..L-1), q[3] sum of the final bond dimensions;
 * tmf_canon_dims: final bond dimensions and charges (before the `cutoff` truncation, which the caller applies to
 * the returned singular values: a discarded direction only multiplies zeros afterwards).  tmf_canon_run enqueues
 * the whole sweep (three launches per site and sweep, no host synchronisation): T0_dev + t0_off[j] is the tensor of
 * site j; outputs: right-canonical tensors (site j at sum_{i<j} dims2[i] * 2 * dims2[i+1]), singular values
 * (bond j at sum_{i<j} dims2[i], decreasing inside a sector, un-normalised) and inv_dev[j] = 1 / their norm. */
typedef struct tmf_canon tmf_canon;
tmf_canon *tmf_canon_create(int L, const int *dims0, const int *charges0, const int *qp);
void tmf_canon_destroy(tmf_canon *c);
int tmf_canon_sizes(const tmf_canon *c, int64_t *q);
int tmf_canon_dims(const tmf_canon *c, int *dims2, int *charges2);
int tmf_canon_run(tmf_canon *c, const double *T0_dev, const int64_t *t0_off, void *work_dev, int64_t work_bytes,
                  double *T2_dev, double *S_dev, double *inv_dev, void *stream);

/* K15 / f2 -- orthogonal Procrustes of the iMPS gauge fixing on the device.  replaces: the per-sector npc.svd and
 * U Vh of iMPS.basis_rotation (iMPS.py:150-184) and the two sums behind its error metrics (:139-147, :186-190).
 * One job per charge sector: C is the overlap block (row-major m x n, row stride ldc, in place in the output of
 * tmf_minors_blocks), sk the n Schmidt values of the ket sector; R (row-major, row stride ldr) receives U Vh of
 * C diag(sk^2), metrics[0] = sum |C sk|^2, metrics[1] = sum |(R - C) sk|^2.  Blocks up to min(m, n) = 160. */
typedef struct tmf_procrustes_job {
  const double *C;
  const double *sk;
  double *R;
  double *metrics;
  int64_t ldc, ldr;
  int m, n, pad_[2];
} tmf_procrustes_job;
int64_t tmf_procrustes_workspace(const tmf_procrustes_job *jobs_host, int njobs);
int tmf_procrustes_blocks(const tmf_procrustes_job *jobs_host, int njobs, void *work_dev, int64_t work_bytes,
                          void *stream);

/* Multi-GPU plumbing of the sharded conversion (temfpy_b200/dist.py; SURVEY 8(e): "gather of per-site tensors").
 * Peer window: a buffer in the destination rank's HBM, exported with tmf_ipc_export (64-byte CUDA IPC handle +
 * offset of dev_ptr inside its allocation) and mapped by the other ranks of the node with tmf_ipc_open (returns the
 * base of the allocation; add the offset).  Passed as `out_dev` of tmf_chain_tensors, the minors kernels store the
 * site tensors straight into the destination GPU over NVLink -- the gather is fused into the producing kernel.
 * Host segments: tmf_host_register pins a (shared-memory) mapping so that tmf_copy_d2h_async moves a shard to the
 * destination process's address space over the GPU's own PCIe link. */
int tmf_ipc_export(const void *dev_ptr, unsigned char *handle64, int64_t *offset_out);
int tmf_ipc_open(const unsigned char *handle64, void **base_out);
int tmf_ipc_close(void *base);
int tmf_host_register(void *ptr, int64_t bytes);
int tmf_host_unregister(void *ptr);
int tmf_copy_d2h_async(void *dst_host, const void *src_dev, int64_t bytes, void *stream);

/* measurement hooks used by bench.py: kernel launch counter (always on) and per-kernel CUDA-event
 * timing (off by default; enabled only in the profiling pass). */
long long tmf_launch_count(int reset);
int tmf_prof_enable(int on);
int tmf_prof_report(char *buf, int cap);
int tmf_prof_timeline(char *buf, int cap);   /* "tag stream start_ms end_ms" per launch */

/* FP64 peak probe used by bench.py for the roofline denominator: runs `iters` dependent-free
 * DFMA chains on every SM and returns the elapsed ms through *ms_out (synchronises). */
int tmf_fp64_peak_probe(int iters, double *sink_dev, float *ms_out, double *flops_out,
                        void *stream);

#ifdef __cplusplus
}
#endif
#endif
