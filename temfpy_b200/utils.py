"""Various utilities (drop-in for ``temfpy.utils``, reference utils.py:8-103)."""
import logging

import numpy as np


def HT(M: np.ndarray) -> np.ndarray:
    """Hermitian conjugate of the input array (utils.py:8-10)."""
    return M.T.conj()


def n_slice(x: slice) -> int:
    """Number of elements returned by a slice, assuming a very long array (utils.py:13-16)."""
    step = x.step or 1
    return (x.stop - x.start) // step


def block_svd(CLR, vL, vR, e, degeneracy_tol: float = 1e-12, overwrite: bool = True):
    """Completes a block singular-value decomposition (utils.py:19-96).

    Host utility kept for API compatibility (tiny k x k problems); inside ``C_to_MPS`` the pairing
    of the centre bond runs in the native chain driver (``tmf_chain_tensors``)."""
    assert vL.shape[1] == vR.shape[1] == e.size, "Mismatched number of eigenvalues and eigenvectors"
    assert vL.shape[0] == CLR.shape[0], "Mismatched row dimension"
    assert vR.shape[0] == CLR.shape[1], "Mismatched column dimension"
    if e.size == 0:
        return vL, vR
    if not overwrite:
        vL, vR = vL.copy(), vR.copy()
    cuts = np.flatnonzero(np.abs(np.diff(e)) > degeneracy_tol) + 1
    for a, b in zip(np.concatenate(([0], cuts)), np.concatenate((cuts, [e.size]))):
        s = HT(vL[:, a:b]) @ CLR @ vR[:, a:b]
        U, _, Vh = np.linalg.svd(s)
        vL[:, a:b] = vL[:, a:b] @ U
        vR[:, a:b] = vR[:, a:b] @ HT(Vh)
    return vL, vR


def normalize_SV(λ: np.ndarray, logger: logging.Logger) -> np.ndarray:
    """Normalises the input array and prints the norm in the logs (utils.py:99-103)."""
    norm = np.linalg.norm(λ)
    logger.info("Norm of Schmidt values: %s", norm)
    return λ / norm
