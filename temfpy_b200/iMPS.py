"""Finite -> infinite MPS for mean-field states (drop-in surface of ``temfpy.iMPS`` for this path).

In scope: the unit-cell conversion behind ``slater.C_to_iMPS`` /
``H_to_iMPS`` (reference slater.py:1356-1565) and the gauge fixing ``basis_rotation`` it calls
(reference iMPS.py:65-192), and :class:`iMPSError`.  The generic TeNPy transfer-matrix path
(``MPS_to_iMPS``, ``overlap_schmidt``; reference iMPS.py:21-62, :233-441) is out of scope (it takes
arbitrary MPS, not mean-field states).

The unit cell has only ``sites_per_cell`` tensors, so instead of the chain object the driver below
sequences the low-level entry points of the C ABI itself: ``tmf_slater_modes_batched`` on both
correlation matrices, ``tmf_slater_pair_bond`` for the two cuts, the native enumeration / planning,
``tmf_site_overlap_schur_batched`` + ``tmf_minors_blocks`` for the cell tensors *and* for the
physical-leg-free gauge overlap <L'_a|L_b> (slater.py:1540), the per-sector SVDs of the orthogonal
Procrustes problem on the host (K15: a handful of small SVDs) and the grouped GEMM for the final
``C . B_0`` contraction (slater.py:1554).
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import logging
import warnings
from typing import NamedTuple

import numpy as np

from . import _lib
from ._lib import check
from .testing import assert_array_less

logger = logging.getLogger(__name__)

_NUMERICAL_TOL = 1e-14   # iMPS.py:16-18
_UNITARY_TOL = 1e-6
_SCHMIDT_TOL = 1e-6


class iMPSError(NamedTuple):
    """Errors introduced during the conversion to iMPS (reference iMPS.py:195-230)."""
    left_unitary: float
    left_schmidt: float
    right_unitary: float
    right_schmidt: float

    @property
    def left_total(self) -> float:
        return (self.left_schmidt ** 2 + self.left_unitary ** 2) ** 0.5

    @property
    def right_total(self) -> float:
        return (self.right_schmidt ** 2 + self.right_unitary ** 2) ** 0.5

    @property
    def total_error(self) -> float:
        return float(np.linalg.norm(self))

    def __repr__(self) -> str:
        fields = [f"    {f}={x:.8e}" for f, x in zip(self._fields, self) if x != 0]
        return "iMPSError()" if not fields else "iMPSError(\n" + (",\n".join(fields)) + "\n)"


def basis_rotation(overlap: np.ndarray, Schmidt_bra, Schmidt_ket, mode: str = "left", *,
                   form: str = "B", numerical_tol=_NUMERICAL_TOL, unitary_tol=_UNITARY_TOL,
                   schmidt_tol=_SCHMIDT_TOL, q_bra=None, q_ket=None):
    """Unitary closest to the overlap of two Schmidt bases (reference iMPS.py:65-192; same positional
    signature) for a dense ``overlap[bra, ket]``: unitarity test, orthogonal Procrustes per charge sector
    (``npc.svd`` is block-wise), Schmidt-mixing test.  The charges of the two legs, which the reference reads
    off the ``npc.Array``, are passed as the keyword-only ``q_bra`` / ``q_ket`` (default: one sector).
    Returns ``(rotation, unitary_error, schmidt_error)``."""
    if q_bra is None:
        q_bra = np.zeros(np.shape(overlap)[0], dtype=np.int64)
    if q_ket is None:
        q_ket = np.zeros(np.shape(overlap)[1], dtype=np.int64)
    mode = mode.lower()
    assert mode in ["left", "right"], f"`mode` must be either 'left' or 'right', got {mode!r}"
    form = form.upper()
    assert form in ["A", "B"], f"`form` must be either 'A' or 'B', got {form!r}"
    Sb, Sk = np.asarray(Schmidt_bra), np.asarray(Schmidt_ket)
    C_Sk = overlap * Sk[None, :]
    err2 = float(np.sum(Sk ** 2) - np.sum(np.abs(C_Sk) ** 2))
    if err2 < 0:
        assert_array_less(abs(err2), numerical_tol,
                          f"{mode.capitalize()} deviation from unitary: the square of the unitary error "
                          f"{err2} is negative and exceeds the numerical tolerance {numerical_tol:.1e}.")
        unitary_error = 0.0
    else:
        unitary_error = float(np.sqrt(err2))
    logger.info("%s deviation from unitary: %.4e", mode.capitalize(), unitary_error)
    if unitary_error > unitary_tol:
        warnings.warn(f"\n{mode.capitalize()} overlap matrix deviates from unitarity by {unitary_error}.\n"
                      "Increasing the bond dimension may be useful.")
    at_cut = (mode, form) in [("left", "A"), ("right", "B")]
    M = C_Sk * Sb[:, None] if at_cut else C_Sk * Sk[None, :]
    R = np.zeros_like(overlap)
    q_bra, q_ket = np.asarray(q_bra), np.asarray(q_ket)
    for q in np.intersect1d(np.unique(q_bra), np.unique(q_ket)):
        r, c = np.flatnonzero(q_bra == q), np.flatnonzero(q_ket == q)
        U, _, Vh = np.linalg.svd(M[np.ix_(r, c)], full_matrices=False)
        R[np.ix_(r, c)] = U @ Vh
    Sb_C = R * Sb[:, None] if at_cut else R * Sk[None, :]
    schmidt_error = float(np.linalg.norm(Sb_C - C_Sk))
    logger.info("%s Schmidt value mixing:   %.4e", mode.capitalize(), schmidt_error)
    if schmidt_error > schmidt_tol:
        warnings.warn(f"\nMixing between unequal Schmidt value sectors on the {mode} side is\n"
                      f"{schmidt_error}. Increasing the number of sites may help.")
    return R, unitary_error, schmidt_error


def _procrustes_errors(err2, schmidt_error, mode, numerical_tol, unitary_tol, schmidt_tol):
    """The bookkeeping of basis_rotation around the two error sums (iMPS.py:139-147, :186-190)."""
    if err2 < 0:
        assert_array_less(abs(err2), numerical_tol,
                          f"{mode.capitalize()} deviation from unitary: the square of the unitary error "
                          f"{err2} is negative and exceeds the numerical tolerance {numerical_tol:.1e}.")
        unitary_error = 0.0
    else:
        unitary_error = float(np.sqrt(err2))
    logger.info("%s deviation from unitary: %.4e", mode.capitalize(), unitary_error)
    if unitary_error > unitary_tol:
        warnings.warn(f"\n{mode.capitalize()} overlap matrix deviates from unitarity by {unitary_error}.\n"
                      "Increasing the bond dimension may be useful.")
    logger.info("%s Schmidt value mixing:   %.4e", mode.capitalize(), schmidt_error)
    if schmidt_error > schmidt_tol:
        warnings.warn(f"\nMixing between unequal Schmidt value sectors on the {mode} side is\n"
                      f"{schmidt_error}. Increasing the number of sites may help.")
    return unitary_error, float(schmidt_error)


def MPS_to_iMPS(*args, **kwargs):
    raise NotImplementedError("MPS_to_iMPS is a generic TeNPy path and out of scope (SURVEY 2.1 #7)")


def overlap_schmidt(*args, **kwargs):
    raise NotImplementedError("overlap_schmidt is a generic TeNPy path and out of scope (SURVEY 2.1 #7)")


# ---------------------------------------------------------------------------------------------
# low-level driver pieces
# ---------------------------------------------------------------------------------------------
class _Modes:
    """Mode extraction of a list of (bond, side) jobs of one correlation matrix."""

    def __init__(self, be, Cd, L, jobs, cutoff, r_sketch=64):
        lib = be.lib
        self.be, self.Cd, self.L, self.jobs = be, Cd, L, jobs
        nj = len(jobs)
        n = np.array([x if s == _lib.SIDE_L else L - x for x, s in jobs], dtype=np.int64)
        self.n = n
        self.v_off = np.concatenate(([0], np.cumsum(n * n)))
        jx = (C.c_int * nj)(*[x for x, _ in jobs])
        js = (C.c_int * nj)(*[s for _, s in jobs])
        voff = (C.c_int64 * nj)(*[int(v) for v in self.v_off[:-1]])
        self.V = be.empty(int(self.v_off[-1]), np.float64)
        ed = be.empty(nj * _lib.TMF_MAX_MODES, np.float64)
        infod = be.empty(nj * 4, np.int32)
        for r in [r for r in (64, 128, 160) if r >= r_sketch]:
            wb = int(lib.tmf_slater_modes_workspace(L, nj, jx, js, r))
            work = be.empty(wb, np.uint8)
            check(lib, lib.tmf_slater_modes_batched(be.ptr(Cd), L, L, nj, jx, js, cutoff, r, voff, be.ptr(self.V),
                                                    be.ptr(ed), be.ptr(infod), be.ptr(work), wb, be.stream))
            be.sync()
            info = be.to_host(infod, nj * 4).reshape(nj, 4)
            if not np.any(info[:, 2] == 1):
                break
        else:
            raise ValueError("entangled spectrum wider than the largest sketch (r_sketch = 160)")
        if np.any(info[:, 2] != 0):
            raise RuntimeError("mode extraction failed")
        self.k = info[:, 0].astype(int)
        self.f = info[:, 1].astype(int)
        self.e = np.ascontiguousarray(be.to_host(ed, nj * _lib.TMF_MAX_MODES).reshape(nj, _lib.TMF_MAX_MODES))

    def vptr(self, j):
        return self.be.ptr(self.V) + 8 * int(self.v_off[j])

    def pair(self, jl, jr, x, deg_tol):
        """block_svd pairing of an LR bond (slater.py:407-410)."""
        be, lib = self.be, self.be.lib
        assert self.k[jl] == self.k[jr], "number of entangled modes differs between the two sides"   # :394
        k = int(self.k[jl])
        if k == 0:
            return
        wb = int(lib.tmf_slater_pair_bond_workspace(self.L, k))
        work = be.empty(wb, np.uint8)
        e = self.e[jl]
        check(lib, lib.tmf_slater_pair_bond(be.ptr(self.Cd), self.L, self.L, x, k,
                                            e.ctypes.data_as(_lib.c_double_p), deg_tol, self.vptr(jl), self.vptr(jr),
                                            be.ptr(work), wb, be.stream))
        be.sync()


class _Bond(NamedTuple):
    k: int
    filled_left: int
    masks: np.ndarray
    lam: np.ndarray
    charge: np.ndarray      # int32, fermion number to the left


def _bond_vectors(lib, trunc, es, ks, fls, L_max):
    nb = len(ks)
    sectors = trunc.sector_list(range(0, L_max + 1))
    if sectors is None:
        sec_p, n_sec = None, -1
    else:
        sec_p, n_sec = (C.c_int * max(len(sectors), 1))(*sectors), len(sectors)
    cap = (trunc.chi_max + 2) if trunc.chi_max is not None else 1 << 14
    S = _lib.TMF_MAX_MODES + 2
    while True:
        e = np.ascontiguousarray(np.stack(es))
        masks = np.zeros((nb, cap), dtype=np.uint64)
        lam = np.zeros((nb, cap))
        charge = np.zeros((nb, cap), dtype=np.int32)
        chi = np.zeros(nb, dtype=np.int32)
        sec_q = np.zeros((nb, S), dtype=np.int32)
        sec_start = np.zeros((nb, S), dtype=np.int32)
        sec_n = np.zeros(nb, dtype=np.int32)
        kk = np.asarray(ks, dtype=np.int32)
        ff = np.asarray(fls, dtype=np.int32)
        ip = lambda a: a.ctypes.data_as(_lib.c_int_p)
        rc = lib.tmf_bond_vectors_batched(nb, e.ctypes.data_as(_lib.c_double_p), ip(kk), ip(ff),
                                          -1 if trunc.chi_max is None else int(trunc.chi_max), float(trunc.svd_min),
                                          float(trunc.degeneracy_tol), sec_p, n_sec, cap,
                                          masks.ctypes.data_as(_lib.c_u64_p), lam.ctypes.data_as(_lib.c_double_p),
                                          ip(charge), ip(chi), ip(sec_q), ip(sec_start), ip(sec_n), 0)
        if rc == -1 and b"capacity" in lib.tmf_last_error() and cap < (1 << 24):
            cap *= 8
            continue
        check(lib, rc)
        break
    return [_Bond(int(ks[b]), int(fls[b]), masks[b, :chi[b]].copy(), lam[b, :chi[b]].copy(),
                  charge[b, :chi[b]].copy()) for b in range(nb)]


class _Plan:
    """tmf_slater_site_plan for one (bra, ket) pair."""

    def __init__(self, lib, mode, n_bra, n_ket, bra: _Bond, f_bra, nferm_bra, ket: _Bond, f_ket, nferm_ket):
        h = _lib.SitePlan()
        chi_b, chi_k = len(bra.lam), len(ket.lam)
        physical = n_bra + 1 == n_ket
        n_rows = 2 * chi_b if physical else chi_b
        self.bra_cols = np.zeros(bra.k + f_bra + 2, dtype=np.int32)
        self.bra_sign = np.zeros(bra.k + f_bra + 2)
        self.ket_cols = np.zeros(ket.k + f_ket + 1, dtype=np.int32)
        self.ket_sign = np.zeros(ket.k + f_ket + 1)
        self.bra_masks = np.zeros(n_rows, dtype=np.uint64)
        self.ket_masks = np.zeros(chi_k, dtype=np.uint64)
        self.row_p = np.zeros(n_rows, dtype=np.int32)
        self.row_alpha = np.zeros(n_rows, dtype=np.int32)
        blocks = np.zeros(6 * (_lib.TMF_MAX_MODES + 4), dtype=np.int32)
        ip = lambda a: a.ctypes.data_as(_lib.c_int_p)
        dp = lambda a: a.ctypes.data_as(_lib.c_double_p)
        up = lambda a: a.ctypes.data_as(_lib.c_u64_p)
        bq = np.ascontiguousarray(bra.charge, dtype=np.int32)
        kq = np.ascontiguousarray(ket.charge, dtype=np.int32)
        check(lib, lib.tmf_slater_site_plan(mode, n_bra, n_ket, bra.k, f_bra, nferm_bra, chi_b, up(bra.masks), ip(bq),
                                            ket.k, f_ket, nferm_ket, chi_k, up(ket.masks), ip(kq), C.byref(h),
                                            ip(self.bra_cols), dp(self.bra_sign), ip(self.ket_cols),
                                            dp(self.ket_sign), up(self.bra_masks), up(self.ket_masks),
                                            ip(self.row_p), ip(self.row_alpha), ip(blocks)))
        self.h = h
        self.blocks = blocks[: 6 * h.n_blocks].reshape(h.n_blocks, 6).copy()
        self.sb0 = h.s_bra - (h.ka_bra - h.k_always)
        self.sk0 = h.s_ket - (h.ka_ket - h.k_always)
        self.rows, self.cols = h.ka_bra + self.sb0, h.ka_ket + self.sk0


def _run_tensors(be, items):
    """items: list of (plan, Vb_ptr, ldb, Vk_ptr, ldk).  Runs overlap/Schur + minors for all of them in
    one launch each and returns, per item, the list of host blocks (row-major n_rows x n_ket)."""
    lib = be.lib
    ns = len(items)
    ints, dbls, u64s = [], [], []
    o_off = s_off = out_off = 0
    meta = []
    for plan, *_ in items:
        h = plan.h
        m = dict(bc=len(ints), kc=len(ints) + plan.rows, bs=len(dbls), ks=len(dbls) + plan.rows,
                 bm=len(u64s), km=len(u64s) + h.n_rows, o=o_off, s=s_off, out=[])
        ints += plan.bra_cols[: plan.rows].tolist() + plan.ket_cols[: plan.cols].tolist()
        dbls += plan.bra_sign[: plan.rows].tolist() + plan.ket_sign[: plan.cols].tolist()
        u64s += plan.bra_masks.tolist() + plan.ket_masks.tolist()
        o_off += plan.rows * plan.cols
        s_off += h.s_bra * h.s_ket
        for b in plan.blocks:
            m["out"].append(out_off)
            out_off += int(b[1]) * int(b[3])
        meta.append(m)
    ints_d = be.from_host(np.array(ints, dtype=np.int32))
    dbls_d = be.from_host(np.array(dbls, dtype=np.float64))
    u64_d = be.from_host(np.array(u64s, dtype=np.uint64).view(np.int64))
    Od, Sd, detd = be.empty(o_off, np.float64), be.empty(s_off, np.float64), be.empty(ns, np.float64)
    outd = be.empty(out_off, np.float64)
    sj = (_lib.SiteJob * ns)()
    nblk = sum(len(p.blocks) for p, *_ in items)
    mb = (_lib.MinorBlock * max(nblk, 1))()
    u = 0
    for i, ((plan, vb, ldb, vk, ldk), m) in enumerate(zip(items, meta)):
        h, j = plan.h, sj[i]
        j.Vb, j.Vk, j.ldb, j.ldk = vb, vk, max(ldb, 1), max(ldk, 1)
        j.bra_cols, j.ket_cols = be.ptr(ints_d) + 4 * m["bc"], be.ptr(ints_d) + 4 * m["kc"]
        j.bra_sign, j.ket_sign = be.ptr(dbls_d) + 8 * m["bs"], be.ptr(dbls_d) + 8 * m["ks"]
        j.O, j.S, j.det = be.ptr(Od) + 8 * m["o"], be.ptr(Sd) + 8 * m["s"], be.ptr(detd) + 8 * i
        j.n_bra, j.n_ket, j.mode, j.physical = h.n_bra, h.n_ket, h.mode, h.physical
        j.ka_bra, j.ka_ket, j.sb, j.sk = h.ka_bra, h.ka_ket, plan.sb0, plan.sk0
        for b, oo in zip(plan.blocks, m["out"]):
            q = mb[u]
            q.S, q.det = j.S, j.det
            q.bra_masks = be.ptr(u64_d) + 8 * (m["bm"] + int(b[0]))
            q.ket_masks = be.ptr(u64_d) + 8 * (m["km"] + int(b[2]))
            q.out = be.ptr(outd) + 8 * oo
            q.s_bra, q.s_ket, q.n_bra, q.n_ket, q.minor = h.s_bra, h.s_ket, int(b[1]), int(b[3]), int(b[4])
            u += 1
    d1 = be.empty(int(lib.tmf_site_desc_bytes(ns)), np.uint8)
    d2 = be.empty(int(lib.tmf_minor_desc_bytes(nblk)), np.uint8)
    check(lib, lib.tmf_site_overlap_schur_batched(sj, ns, be.ptr(d1), be.stream))
    check(lib, lib.tmf_minors_blocks(mb, nblk, be.ptr(d2), be.stream))
    be.sync()
    out = be.to_host(outd, out_off)
    res = []
    for (plan, *_), m in zip(items, meta):
        res.append([out[oo: oo + int(b[1]) * int(b[3])].reshape(int(b[1]), int(b[3]))
                    for b, oo in zip(plan.blocks, m["out"])])
    return res, outd, meta


class CellTensor:
    """Site tensor of the unit cell: dense blocks over charge sectors, ``T[vL, p, vR]``."""

    def __init__(self, chi_L, chi_R, qtotal):
        self.chi_L, self.chi_R, self.qtotal = chi_L, chi_R, qtotal
        self.blocks = []        # (vL index array, p index array, vR index array, values[len(vL) ...])

    def dense(self):
        T = np.zeros((self.chi_L, 2, self.chi_R))
        for vl, p, vr, val in self.blocks:
            T[vl, p, vr] = val
        return T


def slater_C_to_iMPS(C_short, C_long, trunc_par, sites_per_cell, cut, *, diag_tol=1e-8, unitary_tol=_UNITARY_TOL,
                     schmidt_tol=_SCHMIDT_TOL, spinful=None, offset="auto", unit_cell_width=None, as_tenpy=None,
                     _backend=None):
    """reference slater.py:1356-1565 (same parameters); returns ``(BlockMPS(bc="infinite"), iMPSError)``."""
    from . import slater as _sl
    from .mps import BlockMPS
    from .schmidt_utils import to_stopping_condition
    from .utils import normalize_SV
    trunc = to_stopping_condition(trunc_par)
    if unit_cell_width is None:
        unit_cell_width = sites_per_cell
    elif sites_per_cell % unit_cell_width != 0:
        raise ValueError(f"{unit_cell_width = } does not divide {sites_per_cell = }")
    C_short, C_long = np.asarray(C_short), np.asarray(C_long)
    if spinful == "simple":                                                                # :1456-1466
        offset = 2 * round(np.trace(C_short[:cut, :cut]).real) if offset == "auto" else 2 * offset
        C_short = _sl.spinful_correlation_matrix(C_short, False)
        C_long = _sl.spinful_correlation_matrix(C_long, False)
        sites_per_cell, cut = 2 * sites_per_cell, 2 * cut
    elif spinful == "PH":
        C_short = _sl.spinful_correlation_matrix(C_short, True)
        C_long = _sl.spinful_correlation_matrix(C_long, True)
        sites_per_cell, cut = 2 * sites_per_cell, 2 * cut
    elif spinful is not None:
        raise ValueError(f"`spinful` must be 'simple', 'PH', or `None`, got {spinful!r}")
    Ls, Ll = len(C_short), len(C_long)
    assert C_short.shape == (Ls, Ls), f"Got non-square {C_short.shape} correlation matrix"
    assert C_long.shape == (Ll, Ll), f"Got non-square {C_long.shape} correlation matrix"
    assert Ls + sites_per_cell == Ll, ("The given two MPS must differ by one unit cell, got "
                                      f"{Ll} - {Ls} != {sites_per_cell}")
    assert 0 < cut < Ls, "`cut` must lie inside the short chain"
    if offset == "auto":
        offset = round(np.trace(C_short[:cut, :cut]).real)                                 # :1491
    C_short, C_long = _sl._real_or_raise(C_short, "correlation matrix"), _sl._real_or_raise(C_long, "correlation matrix")
    _sl._check_projector(C_short)
    _sl._check_projector(C_long)
    nf_s, nf_l = int(np.round(np.trace(C_short))), int(np.round(np.trace(C_long)))
    be = _backend or _sl._be()
    lib = be.lib
    cutoff = trunc.svd_min ** 2
    cell = sites_per_cell
    Sd, Ld = be.from_host(C_short.ravel()), be.from_host(C_long.ravel())
    L_, R_ = _lib.SIDE_L, _lib.SIDE_R
    ms = _Modes(be, Sd, Ls, [(cut, L_), (cut, R_)], cutoff)
    ml = _Modes(be, Ld, Ll, [(cut, L_), (cut, R_)] + [(cut + i + 1, R_) for i in range(cell - 1)], cutoff)
    ms.pair(0, 1, cut, trunc.degeneracy_tol)
    ml.pair(0, 1, cut, trunc.degeneracy_tol)
    # Schmidt vectors: short@cut, long@cut, long@cut+1 ... long@cut+cell-1
    es = [ms.e[0], ml.e[0]] + [ml.e[2 + i] for i in range(cell - 1)]
    ks = [ms.k[0], ml.k[0]] + [ml.k[2 + i] for i in range(cell - 1)]
    fls = [ms.f[0], ml.f[0]] + [nf_l - ml.k[2 + i] - ml.f[2 + i] for i in range(cell - 1)]
    bonds = _bond_vectors(lib, trunc, es, ks, fls, Ll)
    b_short, b_long = bonds[0], bonds[1]
    # cell tensors (right mode): ket = previous bond of the long chain, bra = next bond (last: short chain)
    items, info = [], []
    for i in range(cell):
        ket = b_long if i == 0 else bonds[1 + i]
        jk = 1 if i == 0 else 1 + i                    # job index in ml (R side of bond cut + i)
        n_ket = Ll - (cut + i)
        if i == cell - 1:
            bra, vb, f_bra, nfb = b_short, ms.vptr(1), int(ms.f[1]), nf_s
        else:
            bra, vb, f_bra, nfb = bonds[2 + i], ml.vptr(2 + i), int(ml.f[2 + i]), nf_l
        plan = _Plan(lib, 1, n_ket - 1, n_ket, bra, f_bra, nfb, ket, int(ml.f[jk]), nf_l)
        items.append((plan, vb, n_ket - 1, ml.vptr(jk), n_ket))
        info.append((bra, ket))
    # gauge overlap <L'_a(short) | L_b(long)> without physical leg (slater.py:1540)
    gplan = _Plan(lib, 0, cut, cut, b_short, int(ms.f[0]), nf_s, b_long, int(ml.f[0]), nf_l)
    items.append((gplan, ms.vptr(0), cut, ml.vptr(0), cut))
    blocks, outd, meta = _run_tensors(be, items)
    # ---- gauge fixing: orthogonal Procrustes per charge sector (iMPS.py:65-192) -----------------
    chi_s, chi_l = len(b_short.lam), len(b_long.lam)
    import os
    gblocks = [(int(b[0]), int(b[1]), int(b[2]), int(b[3])) for b in gplan.blocks]
    on_device = (not os.environ.get("TMF_HOST_PROCRUSTES") and len(gblocks) > 0
                 and all(min(nr, nc) <= 160 for (_, nr, _, nc) in gblocks))
    R, R_d = None, None
    if on_device:
        # tmf_procrustes_blocks: SVD of C diag(S_ket^2) per sector, U Vh and the sums behind the two error metrics, on
        # the overlap blocks where the minors kernel left them; the rotation stays on the device for C . B_0 below
        R_d = be.from_host(np.zeros(chi_s * chi_l))
        sk_d = be.from_host(np.ascontiguousarray(b_long.lam, dtype=np.float64))
        met_d = be.from_host(np.zeros(2 * len(gblocks)))
        pj = (_lib.ProcrustesJob * len(gblocks))()
        for u, (r0, nr, c0, nc) in enumerate(gblocks):
            a0 = int(gplan.row_alpha[r0])
            assert int(gplan.row_alpha[r0 + nr - 1]) == a0 + nr - 1
            pj[u].C = be.ptr(outd) + 8 * int(meta[-1]["out"][u])
            pj[u].sk = be.ptr(sk_d) + 8 * c0
            pj[u].R = be.ptr(R_d) + 8 * (a0 * chi_l + c0)
            pj[u].metrics = be.ptr(met_d) + 16 * u
            pj[u].ldc, pj[u].ldr, pj[u].m, pj[u].n = nc, chi_l, nr, nc
        wb = int(lib.tmf_procrustes_workspace(pj, len(gblocks)))
        pwork = be.empty(wb, np.uint8)
        check(lib, lib.tmf_procrustes_blocks(pj, len(gblocks), be.ptr(pwork), wb, be.stream))
        be.sync()
        mt = be.to_host(met_d, 2 * len(gblocks)).reshape(-1, 2)
        left_unitary, left_schmidt = _procrustes_errors(float(np.sum(np.asarray(b_long.lam) ** 2) - mt[:, 0].sum()),
                                                        float(np.sqrt(mt[:, 1].sum())), "left", _NUMERICAL_TOL,
                                                        unitary_tol, schmidt_tol)
    else:
        Cov = np.zeros((chi_s, chi_l))
        for b, blk in zip(gplan.blocks, blocks[-1]):
            r0, nr, c0, nc = int(b[0]), int(b[1]), int(b[2]), int(b[3])
            Cov[gplan.row_alpha[r0: r0 + nr][:, None], np.arange(c0, c0 + nc)[None, :]] = blk
        R, left_unitary, left_schmidt = basis_rotation(Cov, b_short.lam, b_long.lam, "left", unitary_tol=unitary_tol,
                                                       schmidt_tol=schmidt_tol, q_bra=b_short.charge,
                                                       q_ket=b_long.charge)
    # ---- assemble; first tensor <- R . B_0 on the device (slater.py:1554) -------------------------
    tensors = []
    for i in range(cell):
        plan = items[i][0]
        bra, ket = info[i]
        chi_L = chi_s if i == 0 else len(ket.lam)
        t = CellTensor(chi_L, len(bra.lam), plan.h.qtotal)
        if i > 0:
            for b, blk in zip(plan.blocks, blocks[i]):
                r0, nr, c0, nc = int(b[0]), int(b[1]), int(b[2]), int(b[3])
                rows = slice(r0, r0 + nr)
                t.blocks.append((np.arange(c0, c0 + nc)[None, :], plan.row_p[rows][:, None],
                                 plan.row_alpha[rows][:, None], blk))
        tensors.append(t)
    plan0 = items[0][0]
    gj, keep = [], []
    rt_chunks, rt_off, o2_off = [], 0, 0
    for bi, b in enumerate(plan0.blocks):
        r0, nr, c0, nc, q = int(b[0]), int(b[1]), int(b[2]), int(b[3]), int(b[5])
        arows = np.flatnonzero(b_short.charge == q)              # short-chain indices of the same charge
        if arows.size == 0:
            continue
        assert int(arows[-1]) - int(arows[0]) + 1 == arows.size
        if R_d is None:
            Rt = np.ascontiguousarray(R[np.ix_(arows, np.arange(c0, c0 + nc))].T)      # (nc x na) row-major
            rt_chunks.append(Rt.ravel())
        gj.append((meta[0]["out"][bi], rt_off, o2_off, nr, nc, arows.size, int(arows[0]), c0))
        keep.append((r0, nr, arows))
        rt_off += nc * arows.size
        o2_off += nr * arows.size
    if gj:
        rt_d = be.from_host(np.concatenate(rt_chunks)) if R_d is None else None
        o2_d = be.empty(o2_off, np.float64)
        jobs = (_lib.GemmJob * len(gj))()
        for u, (ao, bo, oo, m, k, n, a0, c0) in enumerate(gj):
            # row-major out (m x n) = B_0 block (m x k: bra rows x long index) . R^T block (k x n: long x short index)
            # == column-major out^T (n x m) = (R^T)^T . (B_0 block)^T
            if R_d is None:
                jobs[u].A, jobs[u].lda, jobs[u].transA = be.ptr(rt_d) + 8 * bo, n, 0
            else:       # R block in place inside the dense rotation matrix (row-major chi_s x chi_l)
                jobs[u].A, jobs[u].lda, jobs[u].transA = be.ptr(R_d) + 8 * (a0 * chi_l + c0), chi_l, 1
            jobs[u].B, jobs[u].ldb, jobs[u].transB = be.ptr(outd) + 8 * ao, k, 0
            jobs[u].C, jobs[u].ldc = be.ptr(o2_d) + 8 * oo, n
            jobs[u].M, jobs[u].N, jobs[u].K = n, m, k
            jobs[u].alpha, jobs[u].beta = 1.0, 0.0
        desc = be.empty(int(lib.tmf_gemm_desc_bytes(len(gj))), np.uint8)
        check(lib, lib.tmf_gemm_grouped(jobs, len(gj), be.ptr(desc), be.stream))
        be.sync()
        o2 = be.to_host(o2_d, o2_off)
        for (ao, bo, oo, m, k, n, _a0, _c0), (r0, nr, arows) in zip(gj, keep):
            rows = slice(r0, r0 + nr)
            tensors[0].blocks.append((arows[None, :], plan0.row_p[rows][:, None], plan0.row_alpha[rows][:, None],
                                      o2[oo: oo + m * n].reshape(m, n)))
    lam0 = normalize_SV(b_short.lam, logger)
    lams = [lam0] + [normalize_SV(bonds[2 + i].lam, logger) for i in range(cell - 1)] + [lam0]    # :1502, :1513
    charges = [b_short.charge.astype(np.int64) - offset] + \
              [bonds[2 + i].charge.astype(np.int64) - offset for i in range(cell - 1)] + \
              [b_short.charge.astype(np.int64) - offset]
    mps = BlockMPS(L=cell, tensors=tensors, lams=lams, charges=charges, form=["B"] * cell,
                   unit_cell_width=unit_cell_width, ortho_center=None, bc="infinite",
                   meta=dict(offset=offset, chi_long=chi_l, qtotal=[t.qtotal for t in tensors]))
    err = iMPSError(left_unitary, left_schmidt, 0.0, 0.0)
    want = importlib.util.find_spec("tenpy") is not None if as_tenpy is None else bool(as_tenpy)
    return (mps.to_tenpy() if want else mps), err
