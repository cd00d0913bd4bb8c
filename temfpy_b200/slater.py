"""Tools for converting Slater determinants into matrix product states (MPS).

Drop-in for ``temfpy.slater`` (reference slater.py): same entry points and keyword surface
(``correlation_matrix``, ``spinful_correlation_matrix``, ``C_to_MPS``, ``H_to_MPS``,
``C_to_iMPS``, ``H_to_iMPS``, ``SchmidtVectors``), with every floating-point stage running in the
hand-written CUDA kernels of ``csrc/`` behind the C ABI of ``include/temfpy_b200.h``.  There is no
CPU fallback: without the CUDA library or without a GPU every entry point raises ``RuntimeError``.

Return type: TeNPy is not installed in this image, so the drivers return a
:class:`temfpy_b200.mps.BlockMPS` (same leg labels, charges, Schmidt values and form) and convert
it with ``.to_tenpy()`` when ``tenpy`` is importable (``as_tenpy=None`` = automatic).
"""
from __future__ import annotations

import importlib.util
import logging
from dataclasses import dataclass
from typing import Literal

import numpy as np

from . import engine, iMPS as _iMPS
from .mps import BlockMPS
from .schmidt_utils import StoppingCondition, to_stopping_condition
from .testing import _DIAG_TOL
from .utils import HT, normalize_SV

logger = logging.getLogger(__name__)

_backend = None


def _be():
    global _backend
    if _backend is None:
        _backend = engine.TorchBackend()
    return _backend


def _want_tenpy(as_tenpy):
    if as_tenpy is None:
        return importlib.util.find_spec("tenpy") is not None
    return bool(as_tenpy)


def _real_or_complex(M):
    """float64 array, or complex128 if the imaginary part does not vanish (slater.py:1178-1179 drops a ~0 one)."""
    M = np.asarray(M)
    if np.iscomplexobj(M):
        if np.allclose(M.imag, 0.0, rtol=0, atol=1e-14):
            return np.ascontiguousarray(M.real, dtype=np.float64)
        return np.ascontiguousarray(M, dtype=np.complex128)
    return np.ascontiguousarray(M, dtype=np.float64)


def _real_or_raise(M, what):
    M = _real_or_complex(M)
    if np.iscomplexobj(M):
        raise NotImplementedError(f"complex {what}: this entry point is real-only (complex Slater determinants are "
                                  "supported by correlation_matrix / C_to_MPS / H_to_MPS)")
    return M


def embed_complex(C: np.ndarray) -> np.ndarray:
    """Re/im-interleaved real embedding of a complex matrix: ``E[2i+a, 2j+b]`` = ``[[Re, -Im], [Im, Re]]`` of
    ``C[i, j]``.  A Hermitian projector becomes a real symmetric projector of twice the rank; the embedding of a
    vector, ``(Re v_0, Im v_0, Re v_1, ...)``, is the same memory as the interleaved complex vector."""
    L, M = C.shape
    E = np.empty((2 * L, 2 * M))
    E[0::2, 0::2] = C.real
    E[1::2, 1::2] = C.real
    E[1::2, 0::2] = C.imag
    E[0::2, 1::2] = -C.imag
    return E


#### High-level functions ####
def correlation_matrix(H: np.ndarray, N: int | None = None, *, _backend=None) -> tuple[np.ndarray, int]:
    r"""Ground-state correlation matrix of a mean-field Hamiltonian (slater.py:1150-1180).

    The one-off ``eigh(H)`` (K1) stays on LAPACK like in the reference; the rank-N update
    ``C = Phi Phi^T`` (K2) runs on the DMMA GEMM (``tmf_corr_build``)."""
    be = _backend or _be()
    H = _real_or_complex(H)
    e, v = np.linalg.eigh(H)
    if N is None:
        occupied = e < 0
        v = v[:, occupied]
        N = int(occupied.sum())
    else:
        v = v[:, :N]
    L = len(H)
    if N == 0:
        return np.zeros((L, L)), 0
    if np.iscomplexobj(v):
        # complex orbitals: C = Phi Phi^H through the real embedding (emb(Phi) emb(Phi)^T = emb(Phi Phi^H)) on the
        # same DMMA GEMM; C is read back from the even columns of the embedded product
        phi = be.from_host(embed_complex(v).ravel())
        Cd = be.empty(4 * L * L, np.float64)
        engine.check(be.lib, be.lib.tmf_corr_build(be.ptr(phi), 2 * L, 2 * N, 2 * N, be.ptr(Cd), 2 * L, be.stream))
        be.sync()
        E = be.to_host(Cd, 4 * L * L).reshape(2 * L, 2 * L)
        return E[0::2, 0::2] + 1j * E[1::2, 0::2], N
    phi = be.from_host(np.ascontiguousarray(v).ravel())
    Cd = be.empty(L * L, np.float64)
    engine.check(be.lib, be.lib.tmf_corr_build(be.ptr(phi), L, N, N, be.ptr(Cd), L, be.stream))
    be.sync()
    return be.to_host(Cd, L * L).reshape(L, L), N


def spinful_correlation_matrix(C: np.ndarray, ph: bool = True):
    r"""Enlarged correlation matrix for spinful fermions (slater.py:1183-1213)."""
    n, m = C.shape
    assert n == m, f"Got non-square {C.shape} correlation matrix"
    C2 = np.zeros((2 * n, 2 * n), dtype=C.dtype)
    C2[::2, ::2] = C
    C2[1::2, 1::2] = (np.eye(n) - C) if ph else C
    return C2


def _prepare_C(C, spinful):
    if spinful == "simple":
        C = spinful_correlation_matrix(C, False)
    elif spinful == "PH":
        C = spinful_correlation_matrix(C, True)
    elif spinful is not None:
        raise ValueError(f"`spinful` must be 'simple', 'PH', or `None`, got {spinful!r}")
    L = len(C)
    assert C.shape == (L, L), f"Got non-square {C.shape} correlation matrix"
    return _real_or_complex(C)


def _check_projector(C, tol=1e-8, be=None, Cd=None):
    """The mode extraction uses C^2 = C (a Slater determinant); anything else is not an input the
    reference could convert either (its centre-bond assertion eL + eR = 1, slater.py:404).
    Checked on 64 rows of C^2 with the grouped GEMM on the device copy when one is given (a NumPy
    product here would leave a pool of spinning BLAS threads competing with the pipeline's host stages)."""
    L = len(C)
    if be is None or Cd is None:
        dev = np.abs(C @ C - C).max() if L <= 256 else np.abs(C[:64] @ C - C[:64]).max()
    else:
        r = min(64, L)
        T = be.empty(r * L + 1, np.float64)
        desc = be.empty(int(be.lib.tmf_gemm_desc_bytes(1)), np.uint8)
        engine.check(be.lib, be.lib.tmf_projector_defect(be.ptr(Cd), L, L, r, be.ptr(T), be.ptr(desc), be.stream))
        dev = float(be.to_host(T[r * L:], 1)[0])
        if dev != dev:
            raise ValueError("`C` contains NaN")
    if dev > tol:
        raise ValueError(f"`C` is not the correlation matrix of a Slater determinant (max|C^2 - C| = {dev:.2e})")


def _chain_to_mps(res: engine.ChainResult, unit_cell_width) -> BlockMPS:
    L = res.L
    lams, charges = engine.bulk_normalized(res, L, logger)                             # slater.py:1296
    for x in range(L + 1):
        if lams[x] is None:      # (results assembled without shard tables)
            l, c = res.lam_charge(x)
            lams[x], charges[x] = normalize_SV(l, logger), c
    oc = res.ortho_center
    return BlockMPS(L=L, tensors=engine.LazySeq(res.sites, L), lams=lams,
                    charges=charges,
                    form=["A"] * oc + ["B"] * (L - oc), unit_cell_width=unit_cell_width,     # slater.py:1348
                    ortho_center=oc, meta=dict(stats=res.stats, bonds=res.bonds))


def C_to_MPS(C: np.ndarray, trunc_par: dict | StoppingCondition, *, diag_tol: float = _DIAG_TOL,
             ortho_center: int = None, spinful: Literal["simple", "PH", None] = None,
             unit_cell_width: int | None = None, as_tenpy: bool | None = None, _backend=None,
             _keep_device: bool | None = None):
    r"""MPS representation of a Slater determinant from its correlation matrix
    (slater.py:1216-1353; same parameters).

    With ``spinful`` set (Abrikosov-fermion states, whose next stop is ``gutzwiller.abrikosov(_ph)``) the site
    tensors also stay resident in HBM behind the returned object, so that the projection reads them where the
    conversion left them; the device memory is released with the MPS."""
    keep = (spinful is not None) if _keep_device is None else bool(_keep_device)
    trunc_par = to_stopping_condition(trunc_par)
    if unit_cell_width is None:
        unit_cell_width = len(C)
    elif len(C) % unit_cell_width != 0:
        raise ValueError(f"{unit_cell_width = } does not divide system size {len(C)}")
    be = _backend or _be()
    C = _prepare_C(C, spinful)
    L = len(C)
    n_fermion = int(np.round(np.trace(C).real))                              # slater.py:414
    logger.info("Central bond %d", ortho_center or L // 2)
    if np.iscomplexobj(C):
        # complex Slater determinant: the library works on the real embedding (cuts at 2x) and returns complex tensors
        E = embed_complex(C)
        Ed = be.from_host(E.ravel())
        _check_projector(E, be=be, Cd=Ed)
        res = engine.run_chain(be, Ed, 2 * L, L, trunc_par, n_fermion, ortho_center=ortho_center, r_sketch=96,
                               cplx=True, keep_device=keep)
        mps = _chain_to_mps(res, unit_cell_width)
        return mps.to_tenpy() if _want_tenpy(as_tenpy) else mps
    Cd = be.from_host(C.ravel())
    _check_projector(C, be=be, Cd=Cd)
    from . import testing
    if testing.TEST_ACTION != "pass" and 0 < (ortho_center or L // 2) < L:
        # the reference's consistency check of the central bond (slater.py:420-421 -> testing.py:131-177)
        SchmidtModes.from_correlation_matrix(C, ortho_center or L // 2, trunc_par, which="LR", diag_tol=diag_tol,
                                             _backend=be)
    res = engine.run_chain(be, Cd, L, L, trunc_par, n_fermion, ortho_center=ortho_center, keep_device=keep)
    mps = _chain_to_mps(res, unit_cell_width)
    return mps.to_tenpy() if _want_tenpy(as_tenpy) else mps


def H_to_MPS(H: np.ndarray, trunc_par: dict | StoppingCondition, *, diag_tol: float = _DIAG_TOL,
             ortho_center: int = None, spinful: Literal["simple", "PH", None] = None,
             unit_cell_width: int | None = None, as_tenpy: bool | None = None, _backend=None):
    r"""MPS representation of a Slater determinant from its single body Hamiltonian
    (slater.py:1568-1627)."""
    C, _ = correlation_matrix(H, _backend=_backend)
    return C_to_MPS(C, trunc_par, diag_tol=diag_tol, ortho_center=ortho_center, spinful=spinful,
                    unit_cell_width=unit_cell_width, as_tenpy=as_tenpy, _backend=_backend)


#### Schmidt modes of a single bond ####
@dataclass(frozen=True)
class SchmidtModes:
    r"""Schmidt modes of a Slater determinant on one bond (reference slater.py:41-490).

    Same fields as the reference's dataclass.  ``vL`` / ``vR`` hold the orbitals the conversion uses -- the
    filled and the entangled eigenvectors of ``C[:x,:x]`` / ``C[x:,x:]``; the never-occupied ("empty")
    eigenvectors, which the reference computes and then drops (slater.py:792-797), are omitted, so
    ``ixL["empty"]`` / ``ixR["empty"]`` are empty slices and the matrices are ``n x (filled + entangled)``.
    Column order as in the reference: L = filled, entangled (decreasing eigenvalue); R = entangled, filled."""
    e: np.ndarray
    vL: np.ndarray | None
    vR: np.ndarray | None
    ixL: dict | None
    ixR: dict | None
    nL: int
    nR: int
    n_fermion: int

    @property
    def n_entangled(self) -> int:
        return self.e.size

    def size(self, which: str = "T") -> int:
        w = which[0].upper()
        if w not in "LRT":
            raise ValueError("`which` must start with L, R, or T, got " + repr(which))
        return self.nL if w == "L" else self.nR if w == "R" else self.nL + self.nR

    def n_filled(self, which: str) -> int:
        """slater.py:145-174."""
        w = which[0].upper()
        n = lambda sl: sl.stop - sl.start
        if w == "L":
            return n(self.ixL["filled"]) if self.ixL is not None else self.n_fermion - self.n_entangled - n(self.ixR["filled"])
        if w == "R":
            return n(self.ixR["filled"]) if self.ixR is not None else self.n_fermion - self.n_entangled - n(self.ixL["filled"])
        raise ValueError("`which` must start with L or R, got " + repr(which))

    @property
    def vL_entangled(self):
        return None if self.vL is None else self.vL[:, self.ixL["entangled"]]

    @property
    def vR_entangled(self):
        return None if self.vR is None else self.vR[:, self.ixR["entangled"]]

    def mode_vectors(self, which: str, entangled: bool = False):
        w = which[0].upper()
        if w == "L":
            return self.vL_entangled if entangled else self.vL
        if w == "R":
            return self.vR_entangled if entangled else self.vR
        raise ValueError("`which` must start with L or R, got " + which)

    def eigenvalues(self, which: str, entangled: bool = False):
        """slater.py:212-250."""
        w = which[0].upper()
        if w == "L":
            v, ix, e = self.vL, self.ixL, self.e
        elif w == "R":
            v, ix, e = self.vR, self.ixR, 1 - self.e[::-1]
        else:
            raise ValueError("`which` must start with L or R, got " + repr(which))
        if v is None:
            return None
        if entangled:
            return e
        E = np.zeros(v.shape[1])
        E[ix["filled"]] = 1
        E[ix["entangled"]] = e
        return E

    @property
    def singular_values(self):
        """slater.py:252-268."""
        if self.vL is None or self.vR is None:
            return None
        SV = (self.e * (1 - self.e)) ** 0.5
        return SV * (-1) ** (np.arange(SV.size)[::-1])

    @property
    def e_ratio(self):
        """slater.py:425-428."""
        return np.log((1 - self.e) / self.e)

    @classmethod
    def from_correlation_matrix(cls, C, x, trunc_par, *, which="LR", diag_tol=_DIAG_TOL, _backend=None):
        r"""Schmidt modes for a cut between sites ``x-1`` and ``x`` (slater.py:270-423): mode extraction on the
        device (``tmf_slater_modes_batched``), pairing of the two sides of an ``"LR"`` bond
        (``tmf_slater_pair_bond``: ``utils.block_svd`` + the sign flips of :410) and, for ``"LR"``,
        ``testing.check_schmidt_decomposition`` under the global ``TEST_ACTION``."""
        from . import _lib, testing
        which = which.upper()
        assert ("L" in which) or ("R" in which), "`which` must specify at least one of (L)eft or (R)ight"
        trunc_par = to_stopping_condition(trunc_par)
        be = _backend or _be()
        C = _real_or_raise(C, "correlation matrix")
        L = len(C)
        Cd = be.from_host(C.ravel())
        jobs = [(x, s) for s, w in ((_lib.SIDE_L, "L"), (_lib.SIDE_R, "R")) if w in which]
        m = _iMPS._Modes(be, Cd, L, jobs, trunc_par.svd_min ** 2)                                   # :318, :347
        if len(jobs) == 2:
            m.pair(0, 1, x, trunc_par.degeneracy_tol)                                               # :394-410
        out = {}
        for j, (_, side) in enumerate(jobs):
            n, k, f = int(m.n[j]), int(m.k[j]), int(m.f[j])
            V = be.to_host(m.V[int(m.v_off[j]): int(m.v_off[j]) + n * n], n * n).reshape(n, n).T if n else np.zeros((0, 0))
            ent, fil = V[:, :k], V[:, k: k + f]
            if side == _lib.SIDE_L:      # filled, entangled
                out["L"] = (np.hstack([fil, ent]),
                            dict(filled=slice(0, f), entangled=slice(f, f + k), empty=slice(f + k, f + k)))
            else:                         # entangled (the reference's column j is our mode k-1-j), filled
                out["R"] = (np.hstack([ent[:, ::-1], fil]),
                            dict(empty=slice(0, 0), entangled=slice(0, k), filled=slice(k, k + f)))
        k0 = int(m.k[0])
        modes = cls(e=np.array(m.e[0][:k0]), vL=out.get("L", (None, None))[0], vR=out.get("R", (None, None))[0],
                    ixL=out.get("L", (None, None))[1], ixR=out.get("R", (None, None))[1], nL=int(x), nR=int(L - x),
                    n_fermion=int(np.round(np.trace(C))))                                           # :414
        if which == "LR":
            testing.check_schmidt_decomposition(modes, C, diag_tol)                                  # :420-421
        return modes


#### Schmidt vectors of a single bond ####
@dataclass(frozen=True)
class SchmidtVectors:
    r"""Schmidt vectors of a Slater determinant on one bond (reference slater.py:494-755).

    Holds what the reference's object exposes to users -- occupation ``sets`` of the entangled
    modes, ``schmidt_values`` and the charge table ``idx_L`` -- computed by the native path."""
    e: np.ndarray
    sets: np.ndarray
    schmidt_values: np.ndarray
    idx_L: dict
    n_filled_left: int
    nL: int
    nR: int
    n_fermion: int

    @property
    def n_schmidt(self) -> int:
        return len(self.schmidt_values)

    @property
    def n_entangled(self) -> int:
        return self.e.size

    @classmethod
    def from_correlation_matrix(cls, C, x, trunc_par, *, which="LR", diag_tol=_DIAG_TOL):
        which = which.upper()
        assert ("L" in which) or ("R" in which), "`which` must specify at least one of (L)eft or (R)ight"
        trunc_par = to_stopping_condition(trunc_par)
        be = _be()
        C = _real_or_raise(C, "correlation matrix")
        L = len(C)
        nf = int(np.round(np.trace(C)))
        oc = x if which == "LR" else (min(x + 1, L) if which == "L" else max(x - 1, 0))
        lo = min(max(x - 1, 0), L - 1) if which != "R" else min(x, L - 1)
        if which == "R" and x == L:
            lo = L - 1
        Cd = be.from_host(C.ravel())
        res = engine.run_chain(be, Cd, L, L, trunc_par, nf, ortho_center=oc or None, site_lo=lo, site_hi=lo + 1,
                               fetch_tensors=False)
        b = res.bonds[x]
        return cls(e=b.e, sets=b.sets, schmidt_values=b.schmidt_values, idx_L=b.idx_L,
                   n_filled_left=b.filled_left, nL=x, nR=L - x, n_fermion=nf)


def C_to_iMPS(C_short, C_long, trunc_par, sites_per_cell, cut, **kwargs):
    r"""iMPS representation of a Slater determinant from correlation matrices (slater.py:1356-1565)."""
    return _iMPS.slater_C_to_iMPS(C_short, C_long, trunc_par, sites_per_cell, cut, **kwargs)


def H_to_iMPS(H_short, H_long, trunc_par, sites_per_cell, cut, **kwargs):
    r"""iMPS representation of a Slater determinant from Hamiltonians (slater.py:1630-1734)."""
    C_short, _ = correlation_matrix(H_short, _backend=kwargs.get("_backend"))
    C_long, _ = correlation_matrix(H_long, _backend=kwargs.get("_backend"))
    return C_to_iMPS(C_short, C_long, trunc_par, sites_per_cell, cut, **kwargs)


#### MPS tensor data ####
MPSTensorData = engine.SiteTensor
"""Counterpart of the reference's ``MPSTensorData`` (slater.py:872-1143): the site tensor as charge blocks in the
layout ``to_npc_array`` writes (``blocks``, ``row_p``, ``row_alpha``, ``qtotal``, ``plan`` with the orbital counts)."""


#### TeNPy bindings (slater.py:30-36), created on first use: TeNPy is optional here ####
def __getattr__(name):
    if name in ("fermion_site", "fermion_leg", "chinfo"):
        try:
            from tenpy import networks
        except ImportError as err:
            raise AttributeError(f"temfpy_b200.slater.{name} needs physics-tenpy (not installed)") from err
        site = networks.site.FermionSite()
        vals = dict(fermion_site=site, fermion_leg=site.leg, chinfo=site.leg.chinfo)
        globals().update(vals)
        return vals[name]
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
