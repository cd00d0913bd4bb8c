r"""Tools for converting Pfaffian (Bogoliubov / BCS) wave functions into matrix product states.

Drop-in for ``temfpy.pfaffian`` (reference pfaffian.py): same entry points and keyword surface
(``correlation_matrix``, ``C_to_MPS``, ``H_to_MPS``, the basis transformations and Nambu checks).
Every O(n^3) stage runs in the CUDA kernels behind the C ABI of ``include/temfpy_b200.h``; there is
no CPU fallback (without the CUDA library or a GPU every entry point raises ``RuntimeError``).

How the path maps onto the kernels ("Pfaffian path")
----------------------------------------------------------------
In the Majorana basis the Nambu correlation matrix is ``C_M = 1/2 + iA`` with ``A`` real antisymmetric
(pfaffian.py:269-273).  Its real representation ``P'`` (re/im interleaved, ``4L x 4L``) is a real
symmetric projector of rank ``2L`` -- formally the correlation matrix of a Slater determinant -- so

* the per-bond ``eigh`` of pfaffian.py:789 runs through the same sketch / Rayleigh-Ritz /
  pivoted-Cholesky kernels as the Slater path (``tmf_slater_modes_batched`` on ``P'``, cuts at ``4x``);
  ``tmf_pfaffian_pair_modes`` then extracts one complex mode per ``J``-invariant plane;
* ``Vr = V1^+ V2`` and the inverse of its ``U^*`` block (pfaffian.py:1339, :1384) become the overlap GEMM
  plus the blocked LU of ``tmf_site_overlap_schur_batched``: the non-entangled ("always") modes are
  eliminated on the device in whatever real basis the Cholesky produced (a Schur complement does not
  depend on the basis of the eliminated block), leaving a matrix of the size of the entangled modes;
* the remaining ``O(k^3)`` algebra (singular values and inverse of the small ``U^*`` block, the antisymmetric
  contraction matrix ``N``, pfaffian.py:1352-1400, with the sign fixes that depend on the vacuum parities) runs
  in ``tmf_pfaffian_site_finish``, one CTA per site, two launches per chain;
* all tensor entries ``Pf(N[idx, idx])`` (pfaffian.py:1429-1479, pfapack in the reference) are computed
  by ``tmf_pfaffians_blocks`` (complex128 Parlett-Reid, one warp per entry).

Vacuum parities (pfaffian.py:396-456, an SVD per bond in the reference) are obtained without extra
``O(n^3)`` work: two neighbouring vacua have opposite parity exactly when their overlap vanishes, i.e.
when the small Schur complement of ``U^*`` is singular; the absolute parities follow from the empty
blocks at the two chain ends.

Return type: :class:`temfpy_b200.mps.BlockMPS` with ``conserve="parity"`` (``.to_tenpy()`` builds the
TeNPy object when ``tenpy`` is importable); ``C_to_iMPS`` / ``H_to_iMPS`` return the unit cell as a
``BlockMPS(bc="infinite")`` plus the ``iMPSError``.  Not supported in this release: more than 16 entangled modes per
bond (``4k <= TMF_MAX_MODES``; raises ``NotImplementedError``).
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import logging
from dataclasses import dataclass, field
from functools import partial

import numpy as np

from . import _lib, engine, iMPS as _iMPS
from ._lib import check
from .mps import BlockMPS
from .schmidt_utils import StoppingCondition, lowest_sums, snap_degenerate, to_stopping_condition
from .testing import _DIAG_TOL, assert_allclose, assert_array_less
from .utils import HT, normalize_SV

logger = logging.getLogger(__name__)

_backend = None


def _be():
    global _backend
    if _backend is None:
        _backend = engine.TorchBackend()
    return _backend


#### BASIS TRANSFORMATIONS BETWEEN COMPLEX FERMION AND MAJORANA BASIS (pfaffian.py:75-184) ####
_C2M = np.array([[1, 1], [1j, -1j]]) / 2 ** 0.5
_M2C = np.array([[1, -1j], [1, 1j]]) / 2 ** 0.5


def _vec(v, M):
    v = np.asarray(v)
    n = v.shape[0]
    assert n % 2 == 0, "Got vector(s) of odd size (cannot be Nambu)"
    w = v.reshape(n // 2, 2, *v.shape[1:])
    return np.einsum("xa...,ca->xc...", w, M).reshape(v.shape)


def vector_C2M(v: np.ndarray) -> np.ndarray:
    r"""Mode vectors from the complex-fermion (:math:`c^\dagger_1, c_1, \dots`) to the Majorana basis."""
    return _vec(v, _C2M)


def vector_M2C(v: np.ndarray) -> np.ndarray:
    r"""Mode vectors from the Majorana to the complex-fermion basis."""
    return _vec(v, _M2C)


def _mat(H, M):
    H = np.asarray(H)
    n, m = H.shape
    assert n % 2 == 0 and m % 2 == 0, "Got a matrix with odd side length (cannot be Nambu)"
    return np.einsum("xayb,ca,db->xcyd", H.reshape(n // 2, 2, m // 2, 2), M, M.conj()).reshape(n, m)


def matrix_C2M(H: np.ndarray) -> np.ndarray:
    r"""Matrix from the complex-fermion to the Majorana basis."""
    return _mat(H, _C2M)


def matrix_M2C(H: np.ndarray) -> np.ndarray:
    r"""Matrix from the Majorana to the complex-fermion basis."""
    return _mat(H, _M2C)


#### UTILITIES FOR NAMBU CORRELATION MATRICES (pfaffian.py:189-299) ####
def assert_nambu(C_: np.ndarray, basis: str = None, offset: float = None, name: str = "", rtol: float = 0,
                 atol: float = 1e-10) -> np.ndarray:
    r"""Checks (strictness: :data:`.testing.TEST_ACTION`) and regularises a Nambu matrix."""
    C_ = np.asarray(C_)
    n, m = C_.shape
    assert n == m > 0, f"Got non-square {name}"
    assert n % 2 == 0, f"Got {name} with odd side length (cannot be Nambu)"
    n //= 2
    tol = dict(atol=atol, rtol=rtol)
    assert_allclose(C_, HT(C_), **tol, err_msg=f"{name} is not Hermitian")
    C_ = (C_ + HT(C_)) / 2
    if basis == "M":
        real = np.eye(2 * n) * offset / 2
        assert_allclose(C_.real, real, **tol, err_msg="Unexpected real parts in Majorana basis")
        C_ = C_.astype(complex)
        C_.real = real
    elif basis == "C":
        err = f"{name.capitalize()} is not Nambu symmetric"
        assert_allclose(C_[::2, ::2], offset * np.eye(n) - C_[1::2, 1::2].conj(), **tol, err_msg=err)
        assert_allclose(C_[1::2, ::2], -C_[::2, 1::2].conj(), **tol, err_msg=err)
        if np.allclose(C_.imag, 0, **tol):
            C_ = C_.real
    elif basis is not None:
        raise ValueError("Invalid `basis` " + repr(basis))
    return C_


assert_nambu_hamiltonian = partial(assert_nambu, offset=0, name="Hamiltonian")
assert_nambu_correlation = partial(assert_nambu, offset=1, name="correlation matrix")


def _embed(CM: np.ndarray) -> np.ndarray:
    """Real representation (re/im interleaved) of a Hermitian matrix in the Majorana basis."""
    n = len(CM)
    P = np.empty((2 * n, 2 * n))
    P[0::2, 0::2] = CM.real
    P[1::2, 1::2] = CM.real
    P[0::2, 1::2] = -CM.imag
    P[1::2, 0::2] = CM.imag
    return P


def correlation_matrix(H: np.ndarray, basis: str | None = None, *, rtol: float = 0, atol: float = 1e-10,
                       _backend=None) -> np.ndarray:
    r"""Ground-state Nambu correlation matrix of a BdG Hamiltonian (pfaffian.py:302-393).

    The one-off ``eigh(H)`` stays on LAPACK like in the reference; the projector ``v v^+`` is built on
    the device in its real representation (``tmf_corr_build``, FP64 tensor-core GEMM)."""
    be = _backend or _be()
    assert basis in [None, "M->M", "M->C", "C->M", "C->C"], \
        f"Invalid basis spec {basis!r}, should be of form '[MC]->[MC]'"
    tol = dict(rtol=rtol, atol=atol)
    H = assert_nambu_hamiltonian(H, None if basis is None else basis[0], **tol)
    n = len(H) // 2
    e, v = np.linalg.eigh(H)
    assert_allclose(e + e[::-1], 0, **tol)
    if np.any(abs(e) < atol):
        raise RuntimeError("Some energy eigenvalues are zero. You need to construct\n"
                           "your own correlation matrix!\n"
                           f"Middle 10 eigenvalues:\n{e[n - 5: n + 5, None]}")
    assert_array_less(e[:n], 0, "Lower half of eigenvalues is not all negative")
    v = v[:, :n]
    if basis == "C->M":
        v = vector_C2M(v)
    elif basis == "M->C":
        v = vector_M2C(v)
    # E(v) = [emb(v_j), J emb(v_j)]  ->  E E^T = real representation of v v^+
    E = np.empty((4 * n, 2 * n))
    E[0::2, 0::2] = v.real
    E[1::2, 0::2] = v.imag
    E[0::2, 1::2] = -v.imag
    E[1::2, 1::2] = v.real
    phi = be.from_host(E.ravel())
    Pd = be.empty(16 * n * n, np.float64)
    check(be.lib, be.lib.tmf_corr_build(be.ptr(phi), 4 * n, 2 * n, 2 * n, be.ptr(Pd), 4 * n, be.stream))
    be.sync()
    P = be.to_host(Pd, 16 * n * n).reshape(4 * n, 4 * n)
    C_ = P[0::2, 0::2] + 1j * P[1::2, 0::2]
    return assert_nambu_correlation(C_, None if basis is None else basis[3], **tol)


#### results ####
@dataclass
class PfBond:
    """Schmidt data of one bond (reference: pfaffian.SchmidtVectors, pfaffian.py:1008-1214)."""
    x: int
    e: np.ndarray                # (k,) entangled eigenvalues <= 1/2, ascending
    sets: np.ndarray             # (chi, k) bool, sorted by (parity, number) of excitations
    schmidt_values: np.ndarray   # un-normalised
    idx_n: dict
    idx_parity: dict
    pL: int = 0
    pR: int = 0

    @property
    def charge(self) -> np.ndarray:
        """Fermion parity to the left of the bond for every Schmidt vector (pfaffian.py:1485-1489)."""
        q = np.zeros(len(self.schmidt_values), dtype=np.int64)
        for par, slc in self.idx_parity.items():
            q[slc] = (par + self.pL) % 2
        return q


@dataclass
class PfSiteTensor:
    """Site tensor of the parity-conserving MPS (reference: MPSTensorData.to_npc_array,
    pfaffian.py:1750-1778): dense complex blocks over (excitation-number) sectors."""
    site: int
    mode: str
    chi_bra: int
    chi_ket: int
    blocks: list = field(default_factory=list)      # (pipe rows (int array), ket slice, ndarray)
    qtotal: int = 0
    norm: float = 1.0

    def dense_pab(self) -> np.ndarray:
        M = np.zeros((2 * self.chi_bra, self.chi_ket), dtype=complex)
        for rows, sk, blk in self.blocks:
            M[rows, sk] = blk
        return M.reshape(2, self.chi_bra, self.chi_ket)      # unsorted pipe: row = p * chi + alpha

    def dense(self) -> np.ndarray:
        T = self.dense_pab()
        return np.transpose(T, (1, 0, 2)) if self.mode == "left" else np.transpose(T, (2, 0, 1))


def _parity_n_argsort(x):
    """pfaffian.py:986-1005."""
    x = np.asarray(x).ravel()
    idx = np.lexsort((np.arange(len(x)), x, x % 2))
    xs = x[idx]

    def bunch(y):
        cuts = np.concatenate(([0], np.flatnonzero(y[1:] != y[:-1]) + 1, [len(y)]))
        return {int(y[cuts[i]]): slice(int(cuts[i]), int(cuts[i + 1])) for i in range(len(cuts) - 1)}
    return idx, bunch(xs), bunch(xs % 2)


def _pack(sets: np.ndarray) -> np.ndarray:
    if sets.shape[1] == 0:
        return np.zeros(len(sets), dtype=np.uint64)
    return (sets.astype(np.uint64) << np.arange(sets.shape[1], dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint64)


def _emb_mat(Q: np.ndarray) -> np.ndarray:
    a, b = Q.shape
    E = np.empty((2 * a, 2 * b))
    E[0::2, 0::2] = Q.real
    E[1::2, 1::2] = Q.real
    E[0::2, 1::2] = -Q.imag
    E[1::2, 0::2] = Q.imag
    return E


def _cplx(M: np.ndarray) -> np.ndarray:
    """J-structured real (2a, 2b) -> complex (a, b)."""
    return M[0::2, 0::2] + 1j * M[1::2, 0::2]


_SINGULAR = 1e-7      # overlap of vacua of opposite parity vanishes identically; the reference itself
                      # refuses genuine overlaps below min_SV = 1e-6 (pfaffian.py:1355-1357)


class _PfChain:
    """One finite-chain conversion on one device."""

    def __init__(self, be, CM, trunc: StoppingCondition, ortho_center, r_sketch=64):
        self.be, self.lib = be, be.lib
        self.tp = trunc
        self.L = len(CM) // 2
        self.oc = ortho_center or self.L // 2
        self.cutoff = trunc.svd_min ** 2
        self.CM = CM
        self.r_sketch = r_sketch
        self._keep = []

    # -- stage 1: modes ---------------------------------------------------------------------------
    def modes(self):
        be, lib, L, oc = self.be, self.lib, self.L, self.oc
        self.Pd = be.from_host(_embed(self.CM).ravel())
        # the kernels use P'^2 = P' (a Bogoliubov vacuum); anything else is not an input the reference converts either
        # (its Nambu / spectrum assertions, pfaffian.py:795-800)
        n4 = 4 * L
        r = min(64, n4)
        T = be.empty(r * n4 + 1, np.float64)
        desc = be.empty(int(lib.tmf_gemm_desc_bytes(1)), np.uint8)
        check(lib, lib.tmf_projector_defect(be.ptr(self.Pd), n4, n4, r, be.ptr(T), be.ptr(desc), be.stream))
        be.sync()
        defect = float(be.to_host(T[r * n4:], 1)[0])
        if not defect <= 1e-8:
            raise ValueError("`C` is not the correlation matrix of a Bogoliubov vacuum "
                             f"(max|C^2 - C| = {defect:.2e} in the Majorana basis)")
        jobs = [(x, _lib.SIDE_L) for x in range(1, oc + 1)] + [(x, _lib.SIDE_R) for x in range(oc, L)]
        self.jobs = jobs
        self.job_of = {j: i for i, j in enumerate(jobs)}
        nj = len(jobs)
        nsites = np.array([x if s == _lib.SIDE_L else L - x for x, s in jobs], dtype=np.int64)
        rows = 4 * nsites
        v_off = np.concatenate(([0], np.cumsum(rows * rows)))
        self.rows, self.v_off, self.nsites = rows, v_off, nsites
        jx = (C.c_int * nj)(*[4 * x for x, _ in jobs])
        js = (C.c_int * nj)(*[s for _, s in jobs])
        voff_c = (C.c_int64 * nj)(*[int(v) for v in v_off[:-1]])
        self.Vd = be.empty(int(v_off[-1]), np.float64)
        ed = be.empty(nj * _lib.TMF_MAX_MODES, np.float64)
        infod = be.empty(nj * 4, np.int32)
        for r in [r for r in (64, 128, 160) if r >= self.r_sketch]:
            wb = int(lib.tmf_slater_modes_workspace(4 * L, nj, jx, js, r))
            work = be.empty(wb, np.uint8)
            check(lib, lib.tmf_slater_modes_batched(be.ptr(self.Pd), 4 * L, 4 * L, nj, jx, js, self.cutoff, r,
                                                    voff_c, be.ptr(self.Vd), be.ptr(ed), be.ptr(infod),
                                                    be.ptr(work), wb, be.stream))
            be.sync()
            info = be.to_host(infod, nj * 4).reshape(nj, 4)
            if not np.any(info[:, 2] == 1):
                break
        else:
            raise ValueError("entangled spectrum wider than the largest sketch (r_sketch = 160)")
        if np.any(info[:, 2] != 0):
            raise RuntimeError(f"mode extraction failed (status {sorted(set(info[:, 2].tolist()))})")
        k4 = info[:, 0].astype(np.int64)
        self.f = info[:, 1].astype(np.int64)
        if np.any(k4 % 4) or np.any(k4 > _lib.TMF_MAX_MODES):
            raise NotImplementedError("more than 16 entangled modes per bond or an unpaired spectrum: "
                                      "not supported by the embedded mode extraction of this release")
        self.k = (k4 // 4).astype(np.int64)
        if np.any(self.f != 2 * (nsites - self.k)):
            raise AssertionError("Entangled modes asymmetrical in spectrum")                # pfaffian.py:821
        # complex modes out of the J-invariant planes
        tmp_off = np.concatenate(([0], np.cumsum([int(lib.tmf_pair_tmp_doubles(int(r))) for r in rows])))
        tmp = be.empty(int(tmp_off[-1]), np.float64)
        eod = be.empty(nj * 16, np.float64)
        std = be.empty(2 * nj, np.int32)
        pj = (_lib.PairJob * nj)()
        for j in range(nj):
            pj[j].V = be.ptr(self.Vd) + 8 * int(v_off[j])
            pj[j].tmp = be.ptr(tmp) + 8 * int(tmp_off[j])
            pj[j].e_raw = be.ptr(ed) + 8 * j * _lib.TMF_MAX_MODES
            pj[j].e_out = be.ptr(eod) + 8 * j * 16
            pj[j].status = be.ptr(std) + 8 * j
            pj[j].kh_out = be.ptr(std) + 8 * j + 4
            pj[j].rows, pj[j].ld, pj[j].k4, pj[j].side = int(rows[j]), int(rows[j]), int(k4[j]), jobs[j][1]
        desc = be.empty(64 * nj, np.uint8)
        check(lib, lib.tmf_pfaffian_pair_modes(pj, nj, float(self.tp.degeneracy_tol), be.ptr(desc), be.stream))
        be.sync()
        st = be.to_host(std, 2 * nj).reshape(nj, 2)
        if np.any(st[:, 0] == 5):
            raise AssertionError("1/2 eigenvalues asymmetrical in spectrum")                 # pfaffian.py:805
        if np.any(st[:, 0] == 6):
            raise AssertionError("1/2 eigenvectors cannot be made real")                     # pfaffian.py:813
        if np.any(st[:, 0] != 0):
            raise RuntimeError(f"pairing of the Majorana eigenvectors failed (status {sorted(set(st[:, 0].tolist()))})")
        self.kh = st[:, 1].astype(np.int64)
        eo = be.to_host(eod, nj * 16).reshape(nj, 16)
        self.e = [eo[j, : self.k[j]].copy() for j in range(nj)]
        jl, jr = self.job_of.get((oc, _lib.SIDE_L)), self.job_of.get((oc, _lib.SIDE_R))
        if jl is not None and jr is not None and (self.kh[jl] > 0 or self.kh[jr] > 0):
            assert self.kh[jl] == self.kh[jr], "Unequal number of 1/2 modes"               # pfaffian.py:845
            self._centre_half_modes(jl, jr, int(self.kh[jl]))
        self._align_filled()

    def _align_filled(self):
        """Makes the basis of the non-entangled ("filled") space of every job orthogonal to its *final* entangled
        columns.  The mode extraction built that basis as the complement of its raw eigenvectors; ``pair_modes_kernel``
        then replaced the raw columns of eigenvalue 1 - e by the Nambu partners of the columns of eigenvalue e.  For a
        mode whose e is close to the cutoff the two differ by the rounding-level mixing angle theta ~ eps / e of an
        eigenvector with its nearly degenerate neighbours, the filled basis overlaps the partner column by theta and
        the vacuum-overlap norms come out wrong by theta^2 (5e-8 at e = 3e-12).  F <- F - E (E^T F) restores the exact
        complement (the basis of an eliminated block does not matter, its span does); the volume the basis loses,
        sqrt(det(1 - G G^T)) with G = E^T F, is divided out of the per-site determinants (``self.vol``)."""
        be, lib = self.be, self.lib
        nj = len(self.jobs)
        self.vol = np.ones(nj)
        todo = [j for j in range(nj) if self.k[j] > 0 and self.f[j] > 0]
        if not todo:
            return
        g_off = np.concatenate(([0], np.cumsum([4 * int(self.k[j]) * int(self.f[j]) for j in todo])))
        m_off = np.concatenate(([0], np.cumsum([(4 * int(self.k[j])) ** 2 for j in todo])))
        Gd = be.empty(int(g_off[-1]), np.float64)
        Md = be.empty(int(m_off[-1]), np.float64)
        desc = be.empty(int(lib.tmf_gemm_desc_bytes(len(todo))), np.uint8)
        for stage in range(3):
            g = (_lib.GemmJob * len(todo))()
            for u, j in enumerate(todo):
                rows, k4, f = int(self.rows[j]), 4 * int(self.k[j]), int(self.f[j])
                E = be.ptr(self.Vd) + 8 * int(self.v_off[j])
                F = E + 8 * rows * k4
                G = be.ptr(Gd) + 8 * int(g_off[u])
                q = g[u]
                q.alpha, q.beta = 1.0, 0.0
                if stage == 0:      # G (k4 x f) = E^T F
                    q.A, q.lda, q.transA = E, rows, 1
                    q.B, q.ldb, q.transB = F, rows, 0
                    q.C, q.ldc, q.M, q.N, q.K = G, k4, k4, f, rows
                elif stage == 1:    # F -= E G
                    q.A, q.lda, q.transA = E, rows, 0
                    q.B, q.ldb, q.transB = G, k4, 0
                    q.C, q.ldc, q.M, q.N, q.K = F, rows, rows, f, k4
                    q.alpha, q.beta = -1.0, 1.0
                else:               # M (k4 x k4) = G G^T
                    q.A, q.lda, q.transA = G, k4, 0
                    q.B, q.ldb, q.transB = G, k4, 1
                    q.C, q.ldc, q.M, q.N, q.K = be.ptr(Md) + 8 * int(m_off[u]), k4, k4, k4, f
            check(lib, lib.tmf_gemm_grouped(g, len(todo), be.ptr(desc), be.stream))
        be.sync()
        Mh = be.to_host(Md, int(m_off[-1]))
        for u, j in enumerate(todo):
            k4 = 4 * int(self.k[j])
            M = Mh[m_off[u]: m_off[u + 1]].reshape(k4, k4)
            self.vol[j] = float(np.sqrt(max(np.linalg.det(np.eye(k4) - M), 0.0)))

    def _centre_half_modes(self, jl, jr, kh):
        """Eigenvalue-1/2 modes on the central bond (pfaffian.py:857-865, :878-891).  The kernel left, on either
        side, an arbitrary real orthonormal basis r_0 .. r_{2kh-1} of the 1/2 eigenspace packed into complex modes
        w_j = (r_j + i r_{kh+j}) / sqrt(2).  The two bases are paired by the SVD of r_L^T Im(C_LR) r_R (a 2kh x 2kh
        matrix, host) and repacked with the reference's conventions -- left: (r_j + i r_{kh+j}) / sqrt 2, right
        (upper / Nambu-partner column): (r_{kh+j} - i r_j) / sqrt 2 -- so that w_L^+ C_LR conj(w_R) is real
        positive like for the modes paired by ``block_svd``.  (The reference's quasi-random rotation of the
        paired bases, :867-874, acts on both sides alike and is a gauge choice; it is not applied.)"""
        be, oc, L = self.be, self.oc, self.L
        k = int(self.k[jl])
        assert int(self.k[jr]) == k, "Unequal number of entangled modes"
        s2 = 2 ** 0.5

        def real_basis(j):
            rows = int(self.rows[j])
            o = int(self.v_off[j]) + rows * 4 * (k - kh)
            cols = np.array(be.to_host(self.Vd[o: o + rows * 4 * kh], rows * 4 * kh)).reshape(4 * kh, rows)
            u = cols[0::4]                                  # emb(w_j): re / im interleaved
            return np.concatenate([s2 * u[:, 0::2], s2 * u[:, 1::2]]).T, rows, o      # (2 n_side, 2kh)
        RL, rows_l, o_l = real_basis(jl)
        RR, rows_r, o_r = real_basis(jr)
        A = np.ascontiguousarray(self.CM[: 2 * oc, 2 * oc:].imag)
        U, _, Vh = np.linalg.svd(RL.T @ A @ RR)                                           # pfaffian.py:862-865
        RL, RR = RL @ U, RR @ Vh.T

        def pack(re, im, rows):
            out = np.empty((4 * kh, rows))
            for j in range(kh):
                u = np.empty(rows)
                u[0::2], u[1::2] = re[:, j] / s2, im[:, j] / s2
                ju = np.empty(rows)
                ju[0::2], ju[1::2] = -u[1::2], u[0::2]                                    # J emb(w)
                cu = u.copy()
                cu[1::2] *= -1                                                            # emb(conj w)
                jcu = np.empty(rows)
                jcu[0::2], jcu[1::2] = u[1::2], u[0::2]                                   # J emb(conj w)
                out[4 * j: 4 * j + 4] = (u, ju, cu, jcu)
            return out.ravel()
        new_l = pack(RL[:, :kh], RL[:, kh:], rows_l)        # w_L = (r_j + i r_{kh+j}) / sqrt 2
        new_r = pack(RR[:, kh:], RR[:, :kh], rows_r)        # conj(w_R) = (r_{kh+j} - i r_j) / sqrt 2
        for o, new in ((o_l, new_l), (o_r, new_r)):
            d = be.from_host(new)
            self.Vd[o: o + new.size] = d[: new.size]
        be.sync()

    def _job(self, x, side):
        """(job index or None for an empty block, k, f)."""
        j = self.job_of.get((x, side))
        if j is None:
            return None, 0, 0
        return j, int(self.k[j]), int(self.f[j])

    # -- stage 2: Schmidt vectors (host enumeration, native) -------------------------------------
    def enumerate(self):
        L, oc = self.L, self.oc
        self.bonds = {}
        for x in range(L + 1):
            j, k, _ = self._job(x, _lib.SIDE_L if x <= oc else _lib.SIDE_R)
            e = self.e[j] if j is not None else np.zeros(0)
            if x == oc:
                jr, kr, _ = self._job(x, _lib.SIDE_R)
                er = self.e[jr] if jr is not None else np.zeros(0)
                assert k == kr, "Unequal number of entangled modes"                          # pfaffian.py:842
                assert_allclose(e, er, rtol=0, atol=max(self.tp.degeneracy_tol, 1e-13),
                                err_msg="Eigenvalues of C_LL and C_RR do not match")         # :848-849
            a = snap_degenerate(np.log((1 - e) / e) / 2, e)                                  # :925, :1189
            _, sets = lowest_sums(a, self.tp, _lib_override=self.lib)
            if len(sets) == 0:
                raise ValueError("No Schmidt vectors left after filtering by `trunc_par.sectors`!")
            idx, idx_n, idx_par = _parity_n_argsort(sets.sum(axis=1))
            sets = sets[idx]
            lam = np.where(sets, e, 1 - e).prod(axis=1) ** 0.5                               # :979
            self.bonds[x] = PfBond(x=x, e=e, sets=sets, schmidt_values=lam, idx_n=idx_n, idx_parity=idx_par)

    # -- centre pairing (utils.block_svd as called from pfaffian.py:855) --------------------------
    def _centre_pairing(self):
        be, lib, oc, L = self.be, self.lib, self.oc, self.L
        jl, k, _ = self._job(oc, _lib.SIDE_L)
        jr, _, _ = self._job(oc, _lib.SIDE_R)
        self.QL = np.eye(k, dtype=complex)
        self.QRup = np.eye(k, dtype=complex)
        if k == 0:
            return
        nl, nr = 4 * oc, 4 * (L - oc)
        # G_emb = W_L^T  P'[:nl, nl:]  W_R,up  with W_L = [emb(w_a), J emb(w_a)], W_R,up = conj partners
        T = be.empty(nl * 2 * k, np.float64)
        G = be.empty(4 * k * k, np.float64)
        idx_l = be.from_host(np.array([4 * a + t for a in range(k) for t in (0, 1)], dtype=np.int32))
        idx_r = be.from_host(np.array([4 * a + t for a in range(k) for t in (2, 3)], dtype=np.int32))
        g = (_lib.GemmJob * 1)()
        desc = be.empty(int(lib.tmf_gemm_desc_bytes(1)), np.uint8)
        Vl = be.ptr(self.Vd) + 8 * int(self.v_off[jl])
        Vr = be.ptr(self.Vd) + 8 * int(self.v_off[jr])
        # T (nl x 2k) = P'_LR (nl x nr) @ W_R,up ;  P' is symmetric: the NumPy buffer read column-major
        # is P'^T = P', so the block starts at row 0, column nl  -> offset nl * ld
        g[0].A = be.ptr(self.Pd) + 8 * (nl * 4 * L)
        g[0].lda, g[0].transA = 4 * L, 0
        g[0].B, g[0].ldb, g[0].transB, g[0].b_idx = Vr, nr, 0, be.ptr(idx_r)
        g[0].C, g[0].ldc = be.ptr(T), nl
        g[0].M, g[0].N, g[0].K = nl, 2 * k, nr
        g[0].alpha, g[0].beta = 1.0, 0.0
        check(lib, lib.tmf_gemm_grouped(g, 1, be.ptr(desc), be.stream))
        g2 = (_lib.GemmJob * 1)()
        g2[0].A, g2[0].lda, g2[0].transA, g2[0].a_idx = Vl, nl, 1, be.ptr(idx_l)
        g2[0].B, g2[0].ldb, g2[0].transB = be.ptr(T), nl, 0
        g2[0].C, g2[0].ldc = be.ptr(G), 2 * k
        g2[0].M, g2[0].N, g2[0].K = 2 * k, 2 * k, nl
        g2[0].alpha, g2[0].beta = 1.0, 0.0
        desc2 = be.empty(int(lib.tmf_gemm_desc_bytes(1)), np.uint8)
        check(lib, lib.tmf_gemm_grouped(g2, 1, be.ptr(desc2), be.stream))
        be.sync()
        Gc = _cplx(be.to_host(G, 4 * k * k).reshape(2 * k, 2 * k).T)      # column-major -> [row, col]
        e = self.bonds[oc].e
        br = np.flatnonzero(np.abs(np.diff(e)) > self.tp.degeneracy_tol) + 1                 # utils.py:71
        for a, b in zip(np.concatenate(([0], br)), np.concatenate((br, [k]))):
            if e[a] >= 0.5 - self.tp.degeneracy_tol:        # 1/2 modes: paired by _centre_half_modes (pfaffian.py:857-865)
                continue
            U, _, Vh = np.linalg.svd(Gc[a:b, a:b])
            self.QL[a:b, a:b] = U
            self.QRup[a:b, a:b] = Vh.conj().T

    # -- stage 3: overlap + elimination of the non-entangled modes (device) -----------------------
    def site_stage(self):
        be, lib, L, oc = self.be, self.lib, self.L, self.oc
        self._centre_pairing()
        sites = []
        ints, o_elems, s_elems = [], 0, 0
        for i in range(L):
            mode = 1 if i >= oc else 0
            side = _lib.SIDE_R if mode else _lib.SIDE_L
            xb, xk = (i + 1, i) if mode else (i, i + 1)
            jb, k1, f1 = self._job(xb, side)
            jk, k2, f2 = self._job(xk, side)
            nb = (L - xb) if mode else xb
            act_b = list(range(k1 - 1, -1, -1)) if mode else list(range(k1))     # reference order of the
            act_k = list(range(k2 - 1, -1, -1)) if mode else list(range(k2))     # active (entangled) modes
            ent_up = [4 * a + t for a in act_b for t in (2, 3)]
            ent_lo = [4 * a + t for a in act_b for t in (0, 1)]
            if mode:      # physical mode first (pfaffian.py:1682-1694)
                bra_some = [-3, -4] + ent_up + [-1, -2] + ent_lo
            else:         # physical mode last (:1667-1679)
                bra_some = ent_up + [-3, -4] + ent_lo + [-1, -2]
            bra_cols = [4 * k1 + c for c in range(f1)] + bra_some
            ket_cols = [4 * k2 + c for c in range(f2)] + [4 * a + t for a in act_k for t in (2, 3)] + \
                       [4 * a + t for a in act_k for t in (0, 1)]
            sb, sk = 4 * (k1 + 1), 4 * k2
            sur_b, sur_k = max(f1 - f2, 0), max(f2 - f1, 0)
            st = dict(i=i, mode=mode, xb=xb, xk=xk, jb=jb, jk=jk, k1=k1, k2=k2, f1=f1, f2=f2, nb=nb, sb=sb, sk=sk,
                      sur_b=sur_b, sur_k=sur_k, bra_off=len(ints), ket_off=len(ints) + len(bra_cols),
                      o_off=o_elems, s_off=s_elems)
            ints += bra_cols + ket_cols
            o_elems += (f1 + sb) * (f2 + sk)
            s_elems += (sb + sur_b) * (sk + sur_k)
            sites.append(st)
        self.sites = sites
        ns = len(sites)
        cols_d = be.from_host(np.array(ints, dtype=np.int32))
        ones_d = be.from_host(np.ones(max(max(s["f1"] + s["sb"], s["f2"] + s["sk"]) for s in sites), dtype=np.float64))
        Od = be.empty(o_elems, np.float64)
        Sd = be.empty(s_elems, np.float64)
        detd = be.empty(ns, np.float64)
        sj = (_lib.SiteJob * ns)()
        for u, s in enumerate(sites):
            j = sj[u]
            # an empty bra block (chain end) has no stored modes: any valid pointer will do (K = 0)
            j.Vb = be.ptr(self.Vd) + 8 * int(self.v_off[s["jb"]]) if s["jb"] is not None else be.ptr(self.Vd)
            j.Vk = be.ptr(self.Vd) + 8 * int(self.v_off[s["jk"]])
            j.bra_cols = be.ptr(cols_d) + 4 * s["bra_off"]
            j.ket_cols = be.ptr(cols_d) + 4 * s["ket_off"]
            j.bra_sign = j.ket_sign = be.ptr(ones_d)
            j.O = be.ptr(Od) + 8 * s["o_off"]
            j.S = be.ptr(Sd) + 8 * s["s_off"]
            j.det = be.ptr(detd) + 8 * u
            j.ldb, j.ldk = max(4 * s["nb"], 1), 4 * (s["nb"] + 1)
            j.n_bra, j.n_ket = s["nb"], s["nb"] + 1
            j.mode, j.physical = s["mode"], 1
            j.ka_bra, j.ka_ket = s["f1"], s["f2"]
            j.sb, j.sk = s["sb"], s["sk"]
            j.emb = 1
        desc = be.empty(int(lib.tmf_site_desc_bytes(ns)), np.uint8)
        check(lib, lib.tmf_site_overlap_schur_batched(sj, ns, be.ptr(desc), be.stream))
        be.sync()
        self.Sd = Sd
        self.S_host = be.to_host(Sd, s_elems) if __import__('os').environ.get('TMF_PF_HOST_FINISH') else None
        self.det_host = be.to_host(detd, ns)

    # -- stage 4: small algebra on the entangled modes (host) --------------------------------------
    def _site_matrix(self, s):
        """The Schur complement of site ``s`` as R[bra, ket] with the surplus always-modes moved to the
        front and the centre-bond rotations applied to the ket columns."""
        sb, sk, sur_b, sur_k = s["sb"], s["sk"], s["sur_b"], s["sur_k"]
        nr, nc = sb + sur_b, sk + sur_k
        S = self.S_host[s["s_off"]: s["s_off"] + nr * nc].reshape(nc, nr).T.copy()      # column-major
        if s["mode"] == 1:      # right tensors: surplus after the sometimes orbitals -> move to the front
            S = np.concatenate((S[sb:], S[:sb]), axis=0)
            S = np.concatenate((S[:, sk:], S[:, :sk]), axis=1)
        k2 = s["k2"]
        rot = None
        if s["i"] == self.oc - 1 and s["mode"] == 0 and k2:        # ket = left side of the centre bond
            rot_lo, rot_up = self.QL, self.QL.conj()
            rot = True
        elif s["i"] == self.oc and s["mode"] == 1 and k2:          # ket = right side of the centre bond
            rot_lo, rot_up = self.QRup.conj(), self.QRup
            rot = True
        if rot:
            order = np.arange(k2)[::-1] if s["mode"] == 1 else np.arange(k2)       # stored active order
            inv = np.argsort(order)
            for base, Q in ((sur_k, rot_up), (sur_k + 2 * k2, rot_lo)):
                blk = S[:, base: base + 2 * k2].reshape(nr, k2, 2)[:, inv, :].reshape(nr, 2 * k2)   # e-ascending
                blk = blk @ _emb_mat(Q)
                S[:, base: base + 2 * k2] = blk.reshape(nr, k2, 2)[:, order, :].reshape(nr, 2 * k2)
        return S

    def _contract(self, s, R, sets_bra, sets_ket, fix, u_p, ket_sign):
        """norm and N of pfaffian.py:1258-1410 from the device Schur complement.  Returns ``None``
        when the two vacua do not overlap (opposite parity)."""
        k1, k2, sur_b, sur_k, mode = s["k1"], s["k2"], s["sur_b"], s["sur_k"], s["mode"]
        a1, a2 = k1 + 1, k2
        R = R.copy()
        phys = 0 if mode == 1 else k1                  # position of the physical mode among the bra actives
        up0, lo0 = sur_b, sur_b + 2 * a1               # first row of the upper / lower bra pairs
        rows_phys = [up0 + 2 * phys, up0 + 2 * phys + 1, lo0 + 2 * phys, lo0 + 2 * phys + 1]
        if u_p != 1.0:
            R[rows_phys] *= u_p
        if ket_sign != 1.0:
            R[:, sur_k:] *= ket_sign
        if fix:                                        # pfaffian.py:1708-1719
            if mode == 1:
                oth = [r for r in range(up0, lo0 + 2 * a1) if r not in rows_phys]
                R[oth] *= -1.0
            up_rows = [up0 + 2 * phys, up0 + 2 * phys + 1]
            lo_rows = [lo0 + 2 * phys, lo0 + 2 * phys + 1]
            R[up_rows + lo_rows] = R[lo_rows + up_rows]
        n_up_r, n_up_c = sur_b + 2 * a1, sur_k + 2 * a2
        X = R[:n_up_r, :n_up_c]
        assert X.shape[0] == X.shape[1], "inconsistent mode counts"
        sv = np.linalg.svd(X, compute_uv=False) if X.size else np.ones(0)
        if sv.size and sv.min() < _SINGULAR:
            return None
        Xi = np.linalg.inv(X) if X.size else X
        idx1 = np.flatnonzero(sets_bra.any(axis=0))                                         # :1361-1374
        idx2 = np.flatnonzero(sets_ket.any(axis=0))[::-1]
        pair = lambda base, t: [base + 2 * t, base + 2 * t + 1]
        R1 = [r for t in idx1 for r in pair(lo0, t)]              # lower bra rows of the active modes
        U1 = [r for t in idx1 for r in pair(up0, t)]              # their upper partners (rows of X)
        U2 = [c for t in idx2 for c in pair(sur_k, t)]            # upper ket columns (columns of X)
        C2 = [c for t in idx2 for c in pair(sur_k + 2 * a2, t)]   # lower ket columns
        AA = _cplx(R[np.ix_(R1, range(n_up_c))] @ Xi[:, U1])                                # :1387
        BA = _cplx(Xi[np.ix_(U2, U1)])                                                      # :1389
        BB = _cplx(Xi[U2] @ R[np.ix_(range(n_up_r), C2)])                                   # :1391
        AA = (AA - AA.T) / 2
        BB = (BB - BB.T) / 2
        N = np.block([[BB, BA], [-BA.T, AA]])
        s1, s2 = sets_bra[:, idx1], sets_ket[:, idx2]
        n1 = np.concatenate((np.zeros((len(s1), s2.shape[1]), bool), s1), axis=1)
        n2 = np.concatenate((s2, np.zeros((len(s2), s1.shape[1]), bool)), axis=1)
        return sv.prod(), N, n1, n2

    def _ext_sets(self, x, mode):
        s = self.bonds[x].sets
        s = s[:, ::-1] if mode == 1 else s                                                  # :955-957
        off, on = np.zeros((len(s), 1), bool), np.ones((len(s), 1), bool)
        return np.block([[off, s], [on, s]]) if mode == 1 else np.block([[s, off], [s, on]])

    def _ket_sets(self, x, mode):
        s = self.bonds[x].sets
        return s[:, ::-1] if mode == 1 else s

    # -- stage 4: finish on the entangled modes (device: tmf_pfaffian_site_finish) -----------------------------
    def _centre_rotations(self, s):
        """Real 2 k2 x 2 k2 matrices that apply the centre-bond rotations of block_svd (pfaffian.py:855) to the upper /
        lower ket pairs of a site next to the centre bond, in the stored order of the active modes; None elsewhere."""
        k2 = s["k2"]
        if s["i"] == self.oc - 1 and s["mode"] == 0 and k2:        # ket = left side of the centre bond
            rot_lo, rot_up = self.QL, self.QL.conj()
        elif s["i"] == self.oc and s["mode"] == 1 and k2:          # ket = right side of the centre bond
            rot_lo, rot_up = self.QRup.conj(), self.QRup
        else:
            return None
        order = np.arange(k2)[::-1] if s["mode"] == 1 else np.arange(k2)
        inv = np.argsort(order)
        pairs = lambda idx: np.array([2 * j + t for j in idx for t in (0, 1)])
        eye = np.eye(2 * k2)
        return tuple(np.ascontiguousarray(eye[:, pairs(inv)] @ _emb_mat(Q)[:, pairs(order)]) for Q in (rot_up, rot_lo))

    def _finish_jobs(self, want_n, fix, u_p, ket_sign, masks, n_off):
        be = self.be
        ns = len(self.sites)
        jobs = (_lib.PfSiteJob * max(ns, 1))()
        for u, s in enumerate(self.sites):
            j = jobs[u]
            j.S = be.ptr(self.Sd) + 8 * s["s_off"]
            if s["rot"] is not None:
                j.rot_up = be.ptr(self.rot_d) + 8 * s["rot"][0]
                j.rot_lo = be.ptr(self.rot_d) + 8 * s["rot"][1]
            j.out = be.ptr(self.fin_out) + 16 * u
            j.sb, j.sk, j.sur_b, j.sur_k, j.mode, j.k1, j.k2 = s["sb"], s["sk"], s["sur_b"], s["sur_k"], s["mode"], \
                s["k1"], s["k2"]
            j.want_n, j.fix, j.u_p, j.ket_sign = int(want_n), int(fix[u]), float(u_p[u]), float(ket_sign[u])
            if want_n:
                j.idx1_mask, j.idx2_mask = masks[u]
                j.N = be.ptr(self.Nd) + 8 * n_off[u]
        desc = be.empty(128 * max(ns, 1), np.uint8)
        check(self.lib, self.lib.tmf_pfaffian_site_finish(jobs, ns, be.ptr(desc), be.stream))
        be.sync()
        return be.to_host(self.fin_out, 2 * ns).reshape(ns, 2).copy()

    def tensors(self):
        import os
        if os.environ.get("TMF_PF_HOST_FINISH"):
            return self._tensors_host()
        be, lib, L, oc = self.be, self.lib, self.L, self.oc
        ns = len(self.sites)
        # centre-bond rotations: uploaded once
        rot_chunks, ro = [], 0
        for s in self.sites:
            g = self._centre_rotations(s)
            s["rot"] = None
            if g is not None:
                s["rot"] = (ro, ro + g[0].size)
                rot_chunks += [g[0].ravel(), g[1].ravel()]
                ro += g[0].size + g[1].size
        self.rot_d = be.from_host(np.concatenate(rot_chunks)) if rot_chunks else be.empty(1, np.float64)
        self.fin_out = be.empty(2 * max(ns, 1), np.float64)
        ones, zeros = np.ones(ns), np.zeros(ns, dtype=np.int64)
        # pass 1: singular values of the U* blocks only -- which overlaps vanish decides the vacuum parities
        out1 = self._finish_jobs(False, zeros, ones, ones, None, None)
        if np.any(out1[:, 0] < 0):
            raise AssertionError("inconsistent mode counts")
        flips = [bool(v < _SINGULAR) for v in out1[:, 1]]
        # absolute vacuum parities from the empty blocks at the chain ends
        pL, pR = {0: 0}, {L: 0}
        for i in range(oc):
            pL[i + 1] = pL[i] ^ int(flips[i])
        for i in reversed(range(oc, L)):
            pR[i] = pR[i + 1] ^ int(flips[i])
        total = pL[oc] ^ pR[oc]
        for x in range(oc + 1, L + 1):
            pL[x] = total ^ pR[x]
        for x in range(oc):
            pR[x] = total ^ pL[x]
        for x, b in self.bonds.items():
            b.pL, b.pR = pL[x], pR[x]
        self.total_parity = total
        # pass 2: the fixes that depend on the parities, N of every site written on the device
        fix = np.array([int(f) for f in flips])
        u_p, ket_sign, masks, n_off, per_site = np.ones(ns), np.ones(ns), [], [], []
        no = 0
        for u, (s, fl) in enumerate(zip(self.sites, flips)):
            mode = s["mode"]
            sets_bra = self._ext_sets(s["xb"], mode).copy()
            if fl:
                c = 0 if mode == 1 else -1
                sets_bra[:, c] = ~sets_bra[:, c]
            sets_ket = self._ket_sets(s["xk"], mode)
            if mode == 0 and pL[s["xb"]] == 1:
                u_p[u] = -1.0                                                                # :1665
            if mode == 1 and s["i"] == oc and pL[oc] == 1:
                ket_sign[u] = -1.0                                                           # :915-916
            idx1 = np.flatnonzero(sets_bra.any(axis=0))                                      # :1361-1374
            idx2 = np.flatnonzero(sets_ket.any(axis=0))[::-1]
            masks.append((int(sum(1 << int(t) for t in idx1)), int(sum(1 << int(t) for t in idx2))))
            m = len(idx1) + len(idx2)
            n_off.append(no)
            no += 2 * m * m
            s1, s2 = sets_bra[:, idx1], sets_ket[:, idx2]
            n1 = np.concatenate((np.zeros((len(s1), s2.shape[1]), bool), s1), axis=1)
            n2 = np.concatenate((s2, np.zeros((len(s2), s1.shape[1]), bool)), axis=1)
            per_site.append((m, n1, n2))
        self.Nd = be.empty(max(no, 1), np.float64)
        out2 = self._finish_jobs(True, fix, u_p, ket_sign, masks, n_off)
        if np.any(out2[:, 1] < _SINGULAR):
            raise AssertionError("Boguliubov vacua do not overlap (U nearly singular)")      # :1355-1357
        blocks, site_meta = [], []
        m_chunks = []
        m_off = out_off = 0
        for u, s in enumerate(self.sites):
            m, n1, n2 = per_site[u]
            vol = (self.vol[s["jb"]] if s["jb"] is not None else 1.0) * (self.vol[s["jk"]] if s["jk"] is not None else 1.0)
            norm = (abs(self.det_host[s["i"]]) / vol * out2[u, 0]) ** 0.25                   # :1352, :1359
            leg_idx, idx_n_bra, _ = _parity_n_argsort(n1.sum(axis=1))                        # :1732
            bm, km = _pack(n1[leg_idx]), _pack(n2)
            m_chunks += [bm, km]
            bra_m_off, ket_m_off = m_off, m_off + len(bm)
            m_off += len(bm) + len(km)
            meta = []
            for nb_, sb_ in idx_n_bra.items():
                for nk_, sk_ in self.bonds[s["xk"]].idx_n.items():
                    if (nb_ + nk_) % 2 == 1:                                                # :1768
                        continue
                    nr, nc = sb_.stop - sb_.start, sk_.stop - sk_.start
                    blocks.append((n_off[u], bra_m_off + sb_.start, ket_m_off + sk_.start, out_off, norm, m, nr, nc,
                                   int(nb_), int(nk_)))
                    meta.append((leg_idx[sb_], sk_, out_off, nr, nc))
                    out_off += 2 * nr * nc
            site_meta.append((s, norm, meta))
        Nd = self.Nd
        Md = be.from_host((np.concatenate(m_chunks).astype(np.uint64) if m_chunks else np.zeros(1, np.uint64)).view(np.int64))
        outd = be.empty(out_off, np.float64)
        nbk = len(blocks)
        pb = (_lib.PfBlock * max(nbk, 1))()
        for u, (no, bo, ko, oo, norm, m, nr, nc, n1_, n2_) in enumerate(blocks):
            pb[u].N = be.ptr(Nd) + 8 * no
            pb[u].bra_masks = be.ptr(Md) + 8 * bo
            pb[u].ket_masks = be.ptr(Md) + 8 * ko
            pb[u].out = be.ptr(outd) + 8 * oo
            pb[u].scale = norm
            pb[u].m, pb[u].n_bra, pb[u].n_ket, pb[u].n1, pb[u].n2 = m, nr, nc, n1_, n2_
        desc = be.empty(int(lib.tmf_pf_desc_bytes(nbk)), np.uint8)
        check(lib, lib.tmf_pfaffians_blocks(pb, nbk, be.ptr(desc), be.stream))
        be.sync()
        out = be.to_host(outd, out_off)
        self.n_pfaffians = out_off // 2
        self.site_tensors = []
        for s, norm, meta in site_meta:
            t = PfSiteTensor(site=s["i"], mode="right" if s["mode"] else "left",
                             chi_bra=len(self.bonds[s["xb"]].schmidt_values),
                             chi_ket=len(self.bonds[s["xk"]].schmidt_values), norm=norm)
            for rows, sk_, oo, nr, nc in meta:
                blk = out[oo: oo + 2 * nr * nc]
                t.blocks.append((rows, sk_, (blk[0::2] + 1j * blk[1::2]).reshape(nr, nc)))
            self.site_tensors.append(t)

    def _tensors_host(self):
        """The same finish with host NumPy (cross-check of the device kernel; TMF_PF_HOST_FINISH=1)."""
        be, lib, L, oc = self.be, self.lib, self.L, self.oc
        Rs, flips = [], []
        for s in self.sites:
            R = self._site_matrix(s)
            Rs.append(R)
            r = self._contract(s, R, self._ext_sets(s["xb"], s["mode"]), self._ket_sets(s["xk"], s["mode"]),
                               False, 1.0, 1.0)
            flips.append(r is None)
        # absolute vacuum parities from the empty blocks at the chain ends
        pL, pR = {0: 0}, {L: 0}
        for i in range(oc):
            pL[i + 1] = pL[i] ^ int(flips[i])
        for i in reversed(range(oc, L)):
            pR[i] = pR[i + 1] ^ int(flips[i])
        total = pL[oc] ^ pR[oc]
        for x in range(oc + 1, L + 1):
            pL[x] = total ^ pR[x]
        for x in range(oc):
            pR[x] = total ^ pL[x]
        for x, b in self.bonds.items():
            b.pL, b.pR = pL[x], pR[x]
        self.total_parity = total
        # contraction matrices, block descriptors
        blocks, site_meta = [], []
        n_chunks, m_chunks = [], []
        n_off = m_off = out_off = 0
        for s, R, fl in zip(self.sites, Rs, flips):
            mode = s["mode"]
            sets_bra = self._ext_sets(s["xb"], mode).copy()
            if fl:
                c = 0 if mode == 1 else -1
                sets_bra[:, c] = ~sets_bra[:, c]
            u_p = -1.0 if (mode == 0 and pL[s["xb"]] == 1) else 1.0                          # :1665
            ket_sign = -1.0 if (mode == 1 and s["i"] == oc and pL[oc] == 1) else 1.0         # :915-916
            res = self._contract(s, R, sets_bra, self._ket_sets(s["xk"], mode), fl, u_p, ket_sign)
            if res is None:
                raise AssertionError("Boguliubov vacua do not overlap (U nearly singular)")  # :1355-1357
            svprod, N, n1, n2 = res
            vol = (self.vol[s["jb"]] if s["jb"] is not None else 1.0) * (self.vol[s["jk"]] if s["jk"] is not None else 1.0)
            norm = (abs(self.det_host[s["i"]]) / vol * svprod) ** 0.25                       # :1352, :1359
            leg_idx, idx_n_bra, _ = _parity_n_argsort(n1.sum(axis=1))                        # :1732
            bm, km = _pack(n1[leg_idx]), _pack(n2)
            m = N.shape[0]
            Nf = np.empty(2 * m * m)
            Nf[0::2], Nf[1::2] = N.real.ravel(), N.imag.ravel()
            n_chunks.append(Nf)
            m_chunks += [bm, km]
            bra_m_off, ket_m_off = m_off, m_off + len(bm)
            m_off += len(bm) + len(km)
            meta = []
            for nb_, sb_ in idx_n_bra.items():
                for nk_, sk_ in self.bonds[s["xk"]].idx_n.items():
                    if (nb_ + nk_) % 2 == 1:                                                # :1768
                        continue
                    nr, nc = sb_.stop - sb_.start, sk_.stop - sk_.start
                    blocks.append((n_off, bra_m_off + sb_.start, ket_m_off + sk_.start, out_off, norm, m, nr, nc,
                                   int(nb_), int(nk_)))
                    meta.append((leg_idx[sb_], sk_, out_off, nr, nc))
                    out_off += 2 * nr * nc
            site_meta.append((s, norm, meta))
            n_off += len(Nf)
        Nd = be.from_host(np.concatenate(n_chunks) if n_chunks else np.zeros(1))
        Md = be.from_host((np.concatenate(m_chunks).astype(np.uint64) if m_chunks else np.zeros(1, np.uint64)).view(np.int64))
        outd = be.empty(out_off, np.float64)
        nbk = len(blocks)
        pb = (_lib.PfBlock * max(nbk, 1))()
        for u, (no, bo, ko, oo, norm, m, nr, nc, n1_, n2_) in enumerate(blocks):
            pb[u].N = be.ptr(Nd) + 8 * no
            pb[u].bra_masks = be.ptr(Md) + 8 * bo
            pb[u].ket_masks = be.ptr(Md) + 8 * ko
            pb[u].out = be.ptr(outd) + 8 * oo
            pb[u].scale = norm
            pb[u].m, pb[u].n_bra, pb[u].n_ket, pb[u].n1, pb[u].n2 = m, nr, nc, n1_, n2_
        desc = be.empty(int(lib.tmf_pf_desc_bytes(nbk)), np.uint8)
        check(lib, lib.tmf_pfaffians_blocks(pb, nbk, be.ptr(desc), be.stream))
        be.sync()
        out = be.to_host(outd, out_off)
        self.n_pfaffians = out_off // 2
        self.site_tensors = []
        for s, norm, meta in site_meta:
            t = PfSiteTensor(site=s["i"], mode="right" if s["mode"] else "left",
                             chi_bra=len(self.bonds[s["xb"]].schmidt_values),
                             chi_ket=len(self.bonds[s["xk"]].schmidt_values), norm=norm)
            for rows, sk_, oo, nr, nc in meta:
                blk = out[oo: oo + 2 * nr * nc]
                t.blocks.append((rows, sk_, (blk[0::2] + 1j * blk[1::2]).reshape(nr, nc)))
            self.site_tensors.append(t)

    def run(self):
        self.modes()
        self.enumerate()
        self.site_stage()
        self.tensors()
        return self


def _prepare_CM(C_, basis, cutoff):
    if basis == "C":
        C_ = matrix_C2M(C_)
    elif basis != "M":
        raise ValueError(f"Argument `basis` must be 'M' or 'C', got {basis!r}")
    return assert_nambu_correlation(C_, "M", atol=cutoff)                                    # pfaffian.py:754


def _want_tenpy(as_tenpy):
    if as_tenpy is None:
        return importlib.util.find_spec("tenpy") is not None
    return bool(as_tenpy)


#### High-level functions ####
def C_to_MPS(C_: np.ndarray, trunc_par: dict | StoppingCondition, *, basis: str, diag_tol: float = _DIAG_TOL,
             ortho_center: int = None, unit_cell_width: int | None = None, as_tenpy: bool | None = None,
             _backend=None):
    r"""MPS representation of a Pfaffian state from its Nambu correlation matrix
    (pfaffian.py:1785-1921; same parameters)."""
    trunc_par = to_stopping_condition(trunc_par)
    L = len(C_) // 2
    if unit_cell_width is None:
        unit_cell_width = L
    elif L % unit_cell_width != 0:
        raise ValueError(f"{unit_cell_width = } does not divide system size {L}")
    be = _backend or _be()
    CM = _prepare_CM(np.asarray(C_), basis, trunc_par.svd_min ** 2)
    chain = _PfChain(be, CM, trunc_par, ortho_center).run()
    oc = chain.oc
    logger.info("Central bond %d", oc)
    lams = [normalize_SV(chain.bonds[x].schmidt_values, logger) for x in range(L + 1)]
    mps = BlockMPS(L=L, tensors=chain.site_tensors, lams=lams,
                   charges=[chain.bonds[x].charge for x in range(L + 1)],
                   form=["A"] * oc + ["B"] * (L - oc), unit_cell_width=unit_cell_width, ortho_center=oc,
                   conserve="parity", meta=dict(bonds=chain.bonds, total_parity=chain.total_parity,
                                                n_pfaffians=chain.n_pfaffians))
    return mps.to_tenpy() if _want_tenpy(as_tenpy) else mps


def H_to_MPS(H: np.ndarray, trunc_par: dict | StoppingCondition, *, basis: str, diag_tol: float = _DIAG_TOL,
             ortho_center: int = None, unit_cell_width: int | None = None, as_tenpy: bool | None = None,
             _backend=None):
    r"""MPS representation of the ground state of a BdG Hamiltonian (pfaffian.py:2094-2148)."""
    C_ = correlation_matrix(H, basis=f"{basis}->{basis}", _backend=_backend)
    return C_to_MPS(C_, trunc_par, basis=basis, diag_tol=diag_tol, ortho_center=ortho_center,
                    unit_cell_width=unit_cell_width, as_tenpy=as_tenpy, _backend=_backend)


#### iMPS (pfaffian.py:1924-2242) ####
@dataclass
class PfCellTensor:
    """Dense site tensor ``T[vL, p, vR]`` of the unit cell (the gauge-rotated first tensor)."""
    T: np.ndarray
    qtotal: int = 0

    def dense(self):
        return self.T


def _rotated_slot(chain, job, Q, sign):
    """Copy of the V slot of ``job`` whose k entangled complex modes are replaced by ``sign * sum_b w_b Q[b, a]``
    (the centre-bond rotation of block_svd, pfaffian.py:855, and the sign of :915-916 baked into the modes; the chain
    driver applies both lazily, to ket columns of the site matrices)."""
    be = chain.be
    rows, k = int(chain.rows[job]), int(chain.k[job])
    o = int(chain.v_off[job])
    slot = chain.Vd[o: o + rows * rows]
    slot = slot.clone() if hasattr(slot, "clone") else slot.copy()
    if k == 0:
        return slot
    cols = np.array(be.to_host(chain.Vd[o: o + rows * 4 * k], rows * 4 * k)).reshape(4 * k, rows)
    W = (cols[0::4, 0::2] + 1j * cols[0::4, 1::2]).T                # (2 n_side, k) complex modes, e ascending
    W = sign * (W @ Q)
    out = np.empty((4 * k, rows))
    for a in range(k):
        u = np.empty(rows)
        u[0::2], u[1::2] = W[:, a].real, W[:, a].imag
        ju = np.empty(rows)
        ju[0::2], ju[1::2] = -u[1::2], u[0::2]
        cu = u.copy()
        cu[1::2] *= -1
        jcu = np.empty(rows)
        jcu[0::2], jcu[1::2] = u[1::2], u[0::2]
        out[4 * a: 4 * a + 4] = (u, ju, cu, jcu)
    d = be.from_host(out.ravel())
    slot[: out.size] = d[: out.size]
    return slot


def _cross_tensor(be, bra, ket, mode, physical, trunc):
    """One tensor between Schmidt vectors of two different chains (pfaffian.py:1578-1748 as called at :2027 and
    :2051): ``bra`` / ``ket`` = (chain, bond, job, V slot, nb sites on that side).  Same device stages as a chain site --
    overlap + elimination of the non-entangled modes, site finish, Pfaffian blocks -- with the parity fix decided by
    the known vacuum parities of the two chains.  Returns (PfSiteTensor, qtotal)."""
    lib = be.lib
    (cb, xb, jb, Vb, nb), (ck, xk, jk, Vk, nk) = bra, ket
    side = _lib.SIDE_R if mode else _lib.SIDE_L
    k1, f1 = int(cb.k[jb]), int(cb.f[jb])
    k2, f2 = int(ck.k[jk]), int(ck.f[jk])
    assert nk == nb + (1 if physical else 0), "bra and ket sizes do not match"
    act_b = list(range(k1 - 1, -1, -1)) if mode else list(range(k1))
    act_k = list(range(k2 - 1, -1, -1)) if mode else list(range(k2))
    ent_up = [4 * a + t for a in act_b for t in (2, 3)]
    ent_lo = [4 * a + t for a in act_b for t in (0, 1)]
    if not physical:
        bra_some = ent_up + ent_lo
    elif mode:
        bra_some = [-3, -4] + ent_up + [-1, -2] + ent_lo
    else:
        bra_some = ent_up + [-3, -4] + ent_lo + [-1, -2]
    bra_cols = [4 * k1 + c for c in range(f1)] + bra_some
    ket_cols = [4 * k2 + c for c in range(f2)] + [4 * a + t for a in act_k for t in (2, 3)] + \
               [4 * a + t for a in act_k for t in (0, 1)]
    sb, sk = 4 * (k1 + (1 if physical else 0)), 4 * k2
    sur_b, sur_k = max(f1 - f2, 0), max(f2 - f1, 0)
    cols_d = be.from_host(np.array(bra_cols + ket_cols, dtype=np.int32))
    ones_d = be.from_host(np.ones(max(f1 + sb, f2 + sk, 1), dtype=np.float64))
    Od = be.empty(max((f1 + sb) * (f2 + sk), 1), np.float64)
    Sd = be.empty(max((sb + sur_b) * (sk + sur_k), 1), np.float64)
    detd = be.empty(1, np.float64)
    sj = (_lib.SiteJob * 1)()
    j = sj[0]
    j.Vb, j.Vk = be.ptr(Vb), be.ptr(Vk)
    j.bra_cols, j.ket_cols = be.ptr(cols_d), be.ptr(cols_d) + 4 * len(bra_cols)
    j.bra_sign = j.ket_sign = be.ptr(ones_d)
    j.O, j.S, j.det = be.ptr(Od), be.ptr(Sd), be.ptr(detd)
    j.ldb, j.ldk = max(4 * nb, 1), 4 * nk
    j.n_bra, j.n_ket = nb, nk
    j.mode, j.physical = mode, int(physical)
    j.ka_bra, j.ka_ket = f1, f2
    j.sb, j.sk = sb, sk
    j.emb = 1
    desc = be.empty(int(lib.tmf_site_desc_bytes(1)), np.uint8)
    check(lib, lib.tmf_site_overlap_schur_batched(sj, 1, be.ptr(desc), be.stream))
    be.sync()
    det = float(be.to_host(detd, 1)[0])
    # parities (pfaffian.py:1631-1646, :1708-1719)
    Bb, Bk = cb.bonds[xb], ck.bonds[xk]
    par_b, par_k = (Bb.pR, Bk.pR) if mode else (Bb.pL, Bk.pL)
    fix = (par_b % 2) != (par_k % 2)
    # (the same fact read off the matrix itself, as the chain driver does: vacua of opposite parity do not overlap)
    pj0 = (_lib.PfSiteJob * 1)()
    outd0 = be.empty(2, np.float64)
    q0 = pj0[0]
    q0.S, q0.out = be.ptr(Sd), be.ptr(outd0)
    q0.sb, q0.sk, q0.sur_b, q0.sur_k, q0.mode, q0.k1, q0.k2 = sb, sk, sur_b, sur_k, mode, k1, k2
    q0.fix, q0.want_n, q0.no_phys, q0.u_p, q0.ket_sign = 0, 0, int(not physical), 1.0, 1.0
    d0 = be.empty(128, np.uint8)
    check(lib, lib.tmf_pfaffian_site_finish(pj0, 1, be.ptr(d0), be.stream))
    be.sync()
    singular = bool(be.to_host(outd0, 2)[1] < _SINGULAR)
    if singular != fix:
        logger.warning("vacuum parities of the two chains (%d, %d) disagree with the overlap of their vacua; "
                       "following the overlap", par_b, par_k)
        fix = singular
    qtotal = ((Bb.pL + Bb.pR) + (Bk.pL + Bk.pR)) % 2 if mode else 0
    sets_b = Bb.sets[:, ::-1] if mode else Bb.sets
    sets_k = Bk.sets[:, ::-1] if mode else Bk.sets
    if physical:
        off, on = np.zeros((len(sets_b), 1), bool), np.ones((len(sets_b), 1), bool)
        sets_bra = np.block([[off, sets_b], [on, sets_b]]) if mode else np.block([[sets_b, off], [sets_b, on]])
    else:
        sets_bra = sets_b.copy()
    if fix:
        c = 0 if mode == 1 else -1
        sets_bra[:, c] = ~sets_bra[:, c]
    u_p = -1.0 if (physical and mode == 0 and Bb.pL == 1) else 1.0
    idx1 = np.flatnonzero(sets_bra.any(axis=0))
    idx2 = np.flatnonzero(sets_k.any(axis=0))[::-1]
    m = len(idx1) + len(idx2)
    Nd = be.empty(max(2 * m * m, 1), np.float64)
    outd = be.empty(2, np.float64)
    pj = (_lib.PfSiteJob * 1)()
    q = pj[0]
    q.S, q.N, q.out = be.ptr(Sd), be.ptr(Nd), be.ptr(outd)
    q.idx1_mask, q.idx2_mask = int(sum(1 << int(t) for t in idx1)), int(sum(1 << int(t) for t in idx2))
    q.sb, q.sk, q.sur_b, q.sur_k, q.mode, q.k1, q.k2 = sb, sk, sur_b, sur_k, mode, k1, k2
    q.fix, q.want_n, q.no_phys, q.u_p, q.ket_sign = int(fix), 1, int(not physical), u_p, 1.0
    d2 = be.empty(128, np.uint8)
    check(lib, lib.tmf_pfaffian_site_finish(pj, 1, be.ptr(d2), be.stream))
    be.sync()
    fin = be.to_host(outd, 2)
    if fin[0] < 0:
        raise AssertionError("inconsistent mode counts")
    if fin[1] < _SINGULAR:
        raise AssertionError("Boguliubov vacua do not overlap (U nearly singular)")
    norm = (abs(det) / (cb.vol[jb] * ck.vol[jk]) * float(fin[0])) ** 0.25
    s1, s2 = sets_bra[:, idx1], sets_k[:, idx2]
    n1 = np.concatenate((np.zeros((len(s1), s2.shape[1]), bool), s1), axis=1)
    n2 = np.concatenate((s2, np.zeros((len(s2), s1.shape[1]), bool)), axis=1)
    leg_idx, idx_n_bra, _ = _parity_n_argsort(n1.sum(axis=1))
    bm, km = _pack(n1[leg_idx]), _pack(n2)
    Md = be.from_host(np.concatenate([bm, km]).astype(np.uint64).view(np.int64))
    blocks, out_off = [], 0
    for nb_, sb_ in idx_n_bra.items():
        for nk_, sk_ in Bk.idx_n.items():
            if (nb_ + nk_) % 2 == 1:
                continue
            nr, nc = sb_.stop - sb_.start, sk_.stop - sk_.start
            blocks.append((sb_, sk_, out_off, nr, nc, int(nb_), int(nk_)))
            out_off += 2 * nr * nc
    outb = be.empty(max(out_off, 1), np.float64)
    pb = (_lib.PfBlock * max(len(blocks), 1))()
    for u, (sb_, sk_, oo, nr, nc, n1_, n2_) in enumerate(blocks):
        pb[u].N = be.ptr(Nd)
        pb[u].bra_masks = be.ptr(Md) + 8 * sb_.start
        pb[u].ket_masks = be.ptr(Md) + 8 * (len(bm) + sk_.start)
        pb[u].out = be.ptr(outb) + 8 * oo
        pb[u].scale = norm
        pb[u].m, pb[u].n_bra, pb[u].n_ket, pb[u].n1, pb[u].n2 = m, nr, nc, n1_, n2_
    d3 = be.empty(int(lib.tmf_pf_desc_bytes(len(blocks))), np.uint8)
    check(lib, lib.tmf_pfaffians_blocks(pb, len(blocks), be.ptr(d3), be.stream))
    be.sync()
    out = be.to_host(outb, max(out_off, 1))
    chi_b, chi_k = len(Bb.schmidt_values), len(Bk.schmidt_values)
    t = PfSiteTensor(site=-1, mode="right" if mode else "left", chi_bra=chi_b, chi_ket=chi_k, norm=norm,
                     qtotal=int(qtotal))
    for (sb_, sk_, oo, nr, nc, _a, _b) in blocks:
        blk = out[oo: oo + 2 * nr * nc]
        t.blocks.append((leg_idx[sb_], sk_, (blk[0::2] + 1j * blk[1::2]).reshape(nr, nc)))
    return t


def C_to_iMPS(C_short, C_long, trunc_par, sites_per_cell, cut, *, basis, diag_tol: float = _DIAG_TOL,
              unitary_tol: float = _iMPS._UNITARY_TOL, schmidt_tol: float = _iMPS._SCHMIDT_TOL,
              unit_cell_width: int | None = None, as_tenpy: bool | None = None, _backend=None):
    r"""iMPS representation of a Nambu mean-field state from the correlation matrices of two chains that differ by
    one unit cell (pfaffian.py:1924-2091; same parameters).  Returns ``(BlockMPS(bc="infinite"), iMPSError)``.

    The right-canonical tensors of the additional unit cell are site tensors of the long chain converted with its
    orthogonality centre at ``cut``; the tensor that closes the cell (right environment of the *short* chain,
    :2018-2027) and the gauge overlap of the two left Schmidt bases (:2051) join Schmidt vectors of different chains
    and run through the same device stages one at a time; the gauge fixing is ``iMPS.basis_rotation`` (:2053-2061).
    Both chains are converted completely (the vacuum parities of this implementation come from the chain ends,
    see the module docstring)."""
    trunc_par = to_stopping_condition(trunc_par)
    C_short, C_long = np.asarray(C_short), np.asarray(C_long)
    L_short, L_long = len(C_short) // 2, len(C_long) // 2
    assert C_short.shape == (2 * L_short, 2 * L_short), f"Got correlation matrix of invalid shape {C_short.shape}."
    assert C_long.shape == (2 * L_long, 2 * L_long), f"Got correlation matrix of invalid shape {C_long.shape}."
    assert L_short + sites_per_cell == L_long, ("The given two MPS must differ by one unit cell, got "
                                                f"{L_long} - {L_short} != {sites_per_cell}")
    if unit_cell_width is None:
        unit_cell_width = sites_per_cell
    elif sites_per_cell % unit_cell_width != 0:
        raise ValueError(f"{unit_cell_width = } does not divide {sites_per_cell = }")
    assert 0 < cut < L_short, "`cut` must lie inside the short chain"
    be = _backend or _be()
    cutoff = trunc_par.svd_min ** 2
    cell = sites_per_cell
    long_ = _PfChain(be, _prepare_CM(C_long, basis, cutoff), trunc_par, cut).run()
    short = _PfChain(be, _prepare_CM(C_short, basis, cutoff), trunc_par, cut).run()
    L_, R_ = _lib.SIDE_L, _lib.SIDE_R
    js_l, js_r = short.job_of[(cut, L_)], short.job_of[(cut, R_)]
    jl_l = long_.job_of[(cut, L_)]
    sgn_s = -1.0 if short.bonds[cut].pL == 1 else 1.0                                     # pfaffian.py:915-916
    Vs_r = _rotated_slot(short, js_r, short.QRup.conj(), sgn_s)
    Vs_l = _rotated_slot(short, js_l, short.QL, 1.0)
    Vl_l = _rotated_slot(long_, jl_l, long_.QL, 1.0)
    # cell tensors: sites cut .. cut + cell - 2 of the long chain, then the closing tensor
    tensors = [long_.site_tensors[cut + i] for i in range(cell - 1)]
    xk = cut + cell - 1
    if cell == 1:      # ket = centre bond of the long chain: its right modes carry the centre rotation too
        jk = long_.job_of[(cut, R_)]
        sgn_l = -1.0 if long_.bonds[cut].pL == 1 else 1.0
        Vk = _rotated_slot(long_, jk, long_.QRup.conj(), sgn_l)
    else:
        jk = long_.job_of[(xk, R_)]
        o = int(long_.v_off[jk])
        Vk = long_.Vd[o: o + int(long_.rows[jk]) ** 2]
    nb = L_short - cut
    tensors.append(_cross_tensor(be, (short, cut, js_r, Vs_r, nb), (long_, xk, jk, Vk, nb + 1), 1, True, trunc_par))
    gauge = _cross_tensor(be, (short, cut, js_l, Vs_l, cut), (long_, cut, jl_l, Vl_l, cut), 0, False, trunc_par)
    b_short, b_long = short.bonds[cut], long_.bonds[cut]
    Cov = np.zeros((gauge.chi_bra, gauge.chi_ket), dtype=complex)
    for rows, sk_, blk in gauge.blocks:
        Cov[rows, sk_] = blk
    R, left_unitary, left_schmidt = _iMPS.basis_rotation(Cov, b_short.schmidt_values, b_long.schmidt_values, "left",
                                                         unitary_tol=unitary_tol, schmidt_tol=schmidt_tol,
                                                         q_bra=b_short.charge, q_ket=b_long.charge)
    first = np.tensordot(R, tensors[0].dense(), axes=(1, 0))                             # pfaffian.py:2063
    tensors[0] = PfCellTensor(first, int(getattr(tensors[0], "qtotal", 0)))
    lam0 = normalize_SV(b_short.schmidt_values, logger)
    lams = [lam0] + [normalize_SV(long_.bonds[cut + i + 1].schmidt_values, logger) for i in range(cell - 1)] + [lam0]
    charges = [b_short.charge] + [long_.bonds[cut + i + 1].charge for i in range(cell - 1)] + [b_short.charge]
    mps = BlockMPS(L=cell, tensors=tensors, lams=lams, charges=charges, form=["B"] * cell,
                   unit_cell_width=unit_cell_width, ortho_center=None, bc="infinite", conserve="parity",
                   meta=dict(qtotal=[int(getattr(t, "qtotal", 0)) for t in tensors]))
    err = _iMPS.iMPSError(left_unitary, left_schmidt, 0.0, 0.0)
    return (mps.to_tenpy() if _want_tenpy(as_tenpy) else mps), err


def H_to_iMPS(H_short, H_long, trunc_par, sites_per_cell, cut, *, basis, diag_tol: float = _DIAG_TOL,
              unitary_tol: float = _iMPS._UNITARY_TOL, schmidt_tol: float = _iMPS._SCHMIDT_TOL,
              unit_cell_width: int | None = None, as_tenpy: bool | None = None, _backend=None):
    r"""iMPS representation from BdG Hamiltonians (pfaffian.py:2094-2242)."""
    C_short = correlation_matrix(H_short, basis=f"{basis}->{basis}", _backend=_backend)
    C_long = correlation_matrix(H_long, basis=f"{basis}->{basis}", _backend=_backend)
    return C_to_iMPS(C_short, C_long, trunc_par, sites_per_cell, cut, basis=basis, diag_tol=diag_tol,
                     unitary_tol=unitary_tol, schmidt_tol=schmidt_tol, unit_cell_width=unit_cell_width,
                     as_tenpy=as_tenpy, _backend=_backend)
