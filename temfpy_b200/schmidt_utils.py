"""Utilities for generating the most significant Schmidt states.

Drop-in for ``temfpy.schmidt_utils`` (reference schmidt_utils.py:18-324): same names, arguments
and error behaviour.  The best-first enumeration itself runs in the native library
(``tmf_lowest_sums``, host C++).
"""
from __future__ import annotations

import ctypes as C
import logging
from collections.abc import Callable, Iterable
from dataclasses import dataclass
from numbers import Number

import numpy as np

logger = logging.getLogger(__name__)

_DEFAULT_SVD_MIN = 1e-6     # schmidt_utils.py:14
_DEFAULT_DEG_TOL = 1e-12    # schmidt_utils.py:15


@dataclass(frozen=True)
class StoppingCondition:
    """Describes a stopping condition for enumerating Schmidt states (schmidt_utils.py:18-185)."""

    sectors: Callable[[int], bool] | Iterable[int] | int | None = None
    chi_max: int | None = None
    svd_min: float | None = None
    degeneracy_tol: float | None = None

    def __post_init__(self):
        if self.svd_min is None:
            object.__setattr__(self, "svd_min", _DEFAULT_SVD_MIN)
        if self.degeneracy_tol is None:
            object.__setattr__(self, "degeneracy_tol", _DEFAULT_DEG_TOL)
        if self.sectors is None:
            is_sector = lambda _: True
        elif isinstance(self.sectors, Number):
            is_sector = lambda x: x == self.sectors
        elif isinstance(self.sectors, Iterable):
            is_sector = lambda x: x in self.sectors
        elif isinstance(self.sectors, Callable):
            is_sector = self.sectors
        else:
            raise TypeError(f"Unexpected `sectors` parameter {self.sectors!r}")
        object.__setattr__(self, "is_sector", is_sector)
        assert (self.chi_max is None or self.chi_max > 0), \
            f"`chi_max` must be a positive integer or None, got {self.chi_max!r}"
        assert 0 < self.svd_min < 1, f"`svd_min` must be between 0 and 1, got {self.svd_min!r}"
        assert self.degeneracy_tol > 0, f"`degeneracy_tol` must be positive, got {self.degeneracy_tol!r}"
        object.__setattr__(self, "max_logval", -np.log(self.svd_min) + self.degeneracy_tol)

    def __call__(self, logvals) -> bool:
        """schmidt_utils.py:99-138."""
        logvals = np.asarray(logvals)
        assert logvals.ndim == 1, f"`logvals` must be a 1D array, got {logvals.ndim!r}"
        if self.chi_max is not None and len(logvals) > self.chi_max:
            return False
        if logvals[-1] - logvals[0] > self.max_logval:
            return False
        return True

    def truncate(self, logvals) -> int:
        """schmidt_utils.py:140-185."""
        logvals = np.asarray(logvals)
        assert logvals.ndim == 1, f"`logvals` must be a 1D array, got {logvals.ndim!r}"
        good = np.ones(len(logvals), dtype=bool)
        if self.chi_max is not None:
            good[self.chi_max:] = False
        good &= (logvals - logvals[0]) < -np.log(self.svd_min)
        gap = np.ones(len(logvals), dtype=bool)
        gap[:-1] = (logvals[1:] - logvals[:-1]) > self.degeneracy_tol
        good &= gap
        return int(np.nonzero(good)[0][-1]) + 1

    def sector_list(self, candidates) -> list | None:
        """Allowed charges among ``candidates`` (None = no filter); feeds the native enumerator,
        which cannot call back into a Python predicate."""
        if self.sectors is None:
            return None
        return [int(q) for q in candidates if self.is_sector(int(q))]


def to_stopping_condition(trunc_par) -> StoppingCondition:
    """schmidt_utils.py:188-208."""
    if isinstance(trunc_par, StoppingCondition):
        return trunc_par
    if isinstance(trunc_par, dict):
        return StoppingCondition(**trunc_par)
    raise TypeError(f"Expected a dictionary or a `StoppingCondition` object, got {trunc_par!r}")


def lowest_sums(a, trunc_par: StoppingCondition, *, filled_left=None, filled_right=None, _lib_override=None):
    """Subsets of ``a`` with the lowest sums (schmidt_utils.py:211-324).

    Returns ``(sums, sets)`` with ``sets`` a bool array (n, len(a)).
    """
    from . import _lib
    lib = _lib_override if _lib_override is not None else _lib.load()
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    assert a.ndim == 1, f"`a` must be a 1D array, got {a.ndim!r}"
    k = a.size
    base = float(np.sum(a[a < 0])) if k else 0.0
    lo = (filled_left if filled_left is not None else (filled_right if filled_right is not None else 0))
    sectors = trunc_par.sector_list(range(lo, lo + k + 1))
    if sectors is None:
        sec_p, n_sec = None, -1
    else:
        sec_p, n_sec = (C.c_int * max(len(sectors), 1))(*sectors), len(sectors)
    cap = (trunc_par.chi_max + 2) if trunc_par.chi_max is not None else 1 << 16
    while True:
        sums = np.zeros(cap)
        sets = np.zeros(cap, dtype=np.uint64)
        n, chk = C.c_int(0), C.c_int(0)
        rc = lib.tmf_lowest_sums(a.ctypes.data_as(_lib.c_double_p), k, base,
                                 -1 if trunc_par.chi_max is None else int(trunc_par.chi_max),
                                 float(trunc_par.svd_min), float(trunc_par.degeneracy_tol), sec_p, n_sec,
                                 -1 if filled_left is None else int(filled_left),
                                 -1 if filled_right is None else int(filled_right), cap,
                                 sums.ctypes.data_as(_lib.c_double_p), sets.ctypes.data_as(_lib.c_u64_p),
                                 C.byref(n), C.byref(chk))
        if rc == -1 and b"capacity" in lib.tmf_last_error() and cap < (1 << 26):
            cap *= 8
            continue
        _lib.check(lib, rc)
        break
    logger.info("Checked %d subsets", chk.value)
    logger.info("Kept %d subsets in charge sectors of interest", n.value)
    m = sets[: n.value]
    bits = ((m[:, None] >> np.arange(k, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    return sums[: n.value].copy(), bits.reshape(n.value, k)


def snap_degenerate(a: np.ndarray, e: np.ndarray) -> np.ndarray:
    """Mode weights that are equal within the accuracy of the eigenvalues ``e`` (4e-15 absolute, amplified
    by ``1 / (2 e (1 - e))`` in ``a = log((1 - e) / e) / 2``) are set to their common mean, so that the
    truncation sees symmetry-related Schmidt multiplets as exact degeneracies instead of splitting them by
    rounding noise (same rule as ``tmf::snap_degenerate`` in csrc/hostlogic.cpp, which the Slater chain
    driver applies; see DESIGN.md "Truncation at noise-level degeneracies")."""
    a = np.array(a, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    k = a.size
    if k < 2:
        return a
    w = e * (1.0 - e)
    tol = np.where(w > 0, 4e-15 / (2.0 * np.where(w > 0, w, 1.0)), 0.0)
    order = np.argsort(np.abs(a), kind="stable")
    s0 = 0
    while s0 < k:
        s1 = s0 + 1
        while s1 < k and abs(a[order[s1]]) - abs(a[order[s1 - 1]]) <= max(tol[order[s1]], tol[order[s1 - 1]]):
            s1 += 1
        if s1 - s0 > 1:
            idx = order[s0:s1]
            a[idx] = np.sign(a[idx]) * np.abs(a[idx]).mean()
        s0 = s1
    return a
