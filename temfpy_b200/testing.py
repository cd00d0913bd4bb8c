r"""Numerical tests with controllable strictness (drop-in for ``temfpy.testing``,
reference testing.py:17-177): the global switch :data:`TEST_ACTION` selects ``"raise"``,
``"warn"`` (default) or ``"pass"``."""
import warnings
from typing import Literal

import numpy as np

from .utils import HT

_DIAG_TOL = 1e-8

TEST_ACTION: Literal["raise", "warn", "pass"] = "warn"


class ComparisonWarning(Warning):
    """Generic warning class for failed equality testing or comparison."""


def _shape_mismatch(x, y, strict=False) -> bool:
    if np.ndim(x) == 0 and np.ndim(y) == 0:
        return False
    if np.ndim(x) == 0 or np.ndim(y) == 0:
        return strict
    return np.shape(x) != np.shape(y)


def _dispatch(fn, a, b, err_msg, kwargs, strict):
    if TEST_ACTION == "raise" or _shape_mismatch(a, b, strict):
        fn(a, b, err_msg=err_msg, strict=strict, **kwargs)
    elif TEST_ACTION == "warn":
        try:
            fn(a, b, err_msg="", strict=strict, **kwargs)
        except AssertionError as err:
            warnings.warn("\n" + err_msg + str(err), category=ComparisonWarning)
    elif TEST_ACTION != "pass":
        raise ValueError(f"Invalid value {TEST_ACTION!r} of `temfpy.testing.TEST_ACTION`,\n"
                         "must be one of 'raise', 'warn', 'pass'.")


def assert_allclose(actual, desired, rtol=1e-7, atol=0.0, equal_nan=True, err_msg="", verbose=False, *,
                    strict=False):
    """testing.py:54-93."""
    _dispatch(np.testing.assert_allclose, actual, desired, err_msg,
              dict(rtol=rtol, atol=atol, equal_nan=equal_nan, verbose=verbose), strict)


def assert_array_less(x, y, err_msg="", verbose=False, *, strict=False):
    """testing.py:96-128."""
    _dispatch(np.testing.assert_array_less, x, y, err_msg, dict(verbose=verbose), strict)


def check_schmidt_decomposition(modes, C: np.ndarray, diag_tol: float = _DIAG_TOL):
    """Checks that Schmidt modes and correlation matrix are consistent (testing.py:131-177).

    Works on host copies; only the orbitals the conversion keeps are available (filled and
    entangled), so the filled+entangled part of C_LL / C_RR and C_LR are verified."""
    if TEST_ACTION == "pass":
        return
    tol = dict(rtol=0, atol=diag_tol)
    if modes.vL is not None:
        N = len(modes.vL)
        E = modes.eigenvalues("L")
        assert_allclose(HT(modes.vL) @ modes.vL, np.eye(modes.vL.shape[1]), **tol, err_msg="vL is not unitary")
        CLL = (E * modes.vL) @ HT(modes.vL)
        assert_allclose(CLL, C[:N, :N], **tol, err_msg="vL does not diagonalise C_LL")
    if modes.vR is not None:
        M = len(modes.vR)
        n = len(C) - M
        E = modes.eigenvalues("R")
        assert_allclose(HT(modes.vR) @ modes.vR, np.eye(modes.vR.shape[1]), **tol, err_msg="vR is not unitary")
        CRR = (E * modes.vR) @ HT(modes.vR)
        assert_allclose(CRR, C[n:, n:], **tol, err_msg="vR does not diagonalise C_RR")
    if (modes.vL is not None) and (modes.vR is not None):
        SV = modes.singular_values
        CLR = (SV * modes.vL_entangled) @ HT(modes.vR_entangled[:, ::-1])
        assert_allclose(CLR, C[:N, N:], **tol, err_msg="vL and vR do not SVD C_LR")
