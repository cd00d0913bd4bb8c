"""Finite -> infinite MPS helpers (drop-in surface of ``temfpy.iMPS`` for the mean-field path).

Only the part that the mean-field conversion itself calls is in scope for this release
(SURVEY 8(f) rank 2): ``iMPSError`` and the unit-cell conversion entry used by
``slater.C_to_iMPS``.  The generic TeNPy transfer-matrix path (``MPS_to_iMPS``,
``overlap_schmidt``; reference iMPS.py:21-62, :233-441) is out of scope (not mean-field).
"""
from __future__ import annotations

from dataclasses import dataclass

_UNITARY_TOL = 1e-6   # iMPS.py:16-18
_SCHMIDT_TOL = 1e-6


@dataclass(frozen=True)
class iMPSError:
    """Errors introduced during the conversion to iMPS (reference iMPS.py:195-230)."""
    left_unitary: float
    left_schmidt: float
    right_unitary: float
    right_schmidt: float


def slater_C_to_iMPS(C_short, C_long, trunc_par, sites_per_cell, cut, **kwargs):
    raise NotImplementedError(
        "C_to_iMPS: the unit-cell conversion (reference slater.py:1356-1565) is scheduled after the "
        "finite-chain path (SURVEY 8(f) rank 2); not available in this release")


def MPS_to_iMPS(*args, **kwargs):
    raise NotImplementedError("MPS_to_iMPS is a generic TeNPy path and out of scope (SURVEY 2.1 #7)")


def overlap_schmidt(*args, **kwargs):
    raise NotImplementedError("overlap_schmidt is a generic TeNPy path and out of scope (SURVEY 2.1 #7)")
