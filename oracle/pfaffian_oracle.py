"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Pfaffian (Bogoliubov) state -> MPS path.

A NumPy restatement of the reference algorithm (temfpy/temfpy ``src/temfpy/pfaffian.py``, read-only
checkout at /root/reference; every function cites the reference ``file:line`` it follows).  It is
the checker of the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline leg may import it.  Nothing under ``temfpy_b200/`` does.

Parity pins:
  * the reference ships no tests / golden vectors for this path, and its Pfaffian routine lives in
    a third-party dependency absent from the image (**pfapack**, unpinned in pyproject.toml:37;
    ``pfapack.ctypes.pfaffian`` = Parlett-Reid ``skpfa``; call site pfaffian.py:1425).  The published
    algorithm (Wimmer, ACM TOMS 38, 30 (2012), Alg. "Parlett-Reid") is restated in :func:`pfaffian`;
  * this restatement is pinned (a) against the reference's own code imported in the build container
    through ``oracle/ref_shim.py`` with ``pfaffian.cpf`` bound to :func:`pfaffian` -- fixtures
    ``tests/golden/pfaffian_*.npz`` written by ``oracle/make_golden.py`` -- and (b) against exact
    known answers: the Jordan-Wigner correlators <c_i^+ c_j>, <c_i c_j> of the dense state must
    reproduce the input Nambu correlation matrix (the check of the reference's
    examples/pfaffian.py:28-39), and Pf(A)^2 = det(A).

TeNPy is absent, so ``to_npc_array`` (pfaffian.py:1750-1778) is restated densely: the *unsorted*
``LegPipe([fermion_leg, leg_bra])`` (pfaffian.py:1655-1657) puts the physical index major, i.e. pipe
row = p * chi_bra + alpha; tensors are returned as ``T[vL, p, vR]`` with the parity charge of every
virtual index.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from slater_oracle import DenseMPS, Trunc, block_svd, lowest_sums

_M_C2M = np.array([[1, 1], [1j, -1j]]) / 2 ** 0.5      # pfaffian.py:98
_M_M2C = np.array([[1, -1j], [1, 1j]]) / 2 ** 0.5      # pfaffian.py:126


# --------------------------------------------------------------------------------------------
# basis changes (pfaffian.py:75-184)
# --------------------------------------------------------------------------------------------
def _pairs(v, M):
    n = v.shape[0] // 2
    w = v.reshape(n, 2, *v.shape[1:])
    return np.einsum("xa...,ca->xc...", w, M).reshape(v.shape)


def vector_C2M(v):
    return _pairs(np.asarray(v), _M_C2M)


def vector_M2C(v):
    return _pairs(np.asarray(v), _M_M2C)


def _matrix(H, M):
    n, m = H.shape
    return np.einsum("xayb,ca,db->xcyd", H.reshape(n // 2, 2, m // 2, 2), M, M.conj()).reshape(n, m)


def matrix_C2M(H):
    return _matrix(np.asarray(H), _M_C2M)


def matrix_M2C(H):
    return _matrix(np.asarray(H), _M_M2C)


def regularise_nambu(C, basis, offset):
    """pfaffian.py:189-286 without the assertions: Hermitian part; Majorana basis -> real part is
    exactly offset/2 on the diagonal; complex-fermion basis -> real if numerically real."""
    C = np.asarray(C)
    C = (C + C.conj().T) / 2
    if basis == "M":
        C = C.astype(complex)
        C.real = np.eye(len(C)) * offset / 2
    elif basis == "C":
        if np.allclose(C.imag, 0, rtol=0, atol=1e-10):
            C = C.real
    elif basis is not None:
        raise ValueError("Invalid `basis` " + repr(basis))
    return C


def correlation_matrix(H, basis=None, atol=1e-10):
    """pfaffian.py:302-393: projector on the negative-energy Bogoliubov modes."""
    assert basis in [None, "M->M", "M->C", "C->M", "C->C"]
    H = regularise_nambu(H, None if basis is None else basis[0], 0)
    n = len(H) // 2
    e, v = np.linalg.eigh(H)
    if np.any(abs(e) < atol):                                                      # :372-377
        raise RuntimeError("Some energy eigenvalues are zero")
    v = v[:, :n]
    if basis == "C->M":
        v = vector_C2M(v)
    elif basis == "M->C":
        v = vector_M2C(v)
    C = v @ v.conj().T
    return regularise_nambu(C, None if basis is None else basis[3], 1)


def bdg_chain(L, t=1.0, mu=0.0, delta=0.05, rng=None, disorder=0.0):
    """Nambu Hamiltonian of a p-wave chain in the complex-fermion basis (layout pfaffian.py:343-351;
    BASELINE cfg2: t=1, mu=0, delta=0.05).  ``disorder`` adds random on-site energies."""
    h = np.zeros((L, L))
    d = np.zeros((L, L))
    i = np.arange(L - 1)
    h[i, i + 1] = h[i + 1, i] = -t
    h[np.arange(L), np.arange(L)] = -mu
    if rng is not None and disorder:
        h[np.arange(L), np.arange(L)] += disorder * rng.normal(size=L)
    d[i, i + 1] = delta
    d[i + 1, i] = -delta
    H = np.zeros((2 * L, 2 * L))
    H[::2, ::2] = h
    H[1::2, 1::2] = -h.conj()
    H[::2, 1::2] = d
    H[1::2, ::2] = -d.conj()
    return H


def random_bdg(L, seed, decay=2.0, cplx=True):
    """Random exponentially decaying Nambu Hamiltonian (pattern of examples/pfaffian.py:13-17)."""
    rng = np.random.default_rng(seed)
    dist = np.abs(np.subtract.outer(np.arange(L), np.arange(L)))
    h = rng.normal(size=(L, L)) + (1j * rng.normal(size=(L, L)) if cplx else 0)
    h = (h + h.conj().T) * np.exp(-dist / decay)
    d = rng.normal(size=(L, L)) + (1j * rng.normal(size=(L, L)) if cplx else 0)
    d = (d - d.T) * np.exp(-dist / decay)
    H = np.zeros((2 * L, 2 * L), dtype=complex if cplx else float)
    H[::2, ::2] = h
    H[1::2, 1::2] = -h.conj()
    H[::2, 1::2] = d
    H[1::2, ::2] = -d.conj()
    return H


# --------------------------------------------------------------------------------------------
# vacuum parity (pfaffian.py:396-456)
# --------------------------------------------------------------------------------------------
def vacuum_parity(V, tol=1e-12):
    """Parity of a Bogoliubov vacuum from the singular values of the V block (Bloch-Messiah:
    1,...,1, s1, s1, ..., 0...; the number of exact ones decides)."""
    V = np.asarray(V)
    if len(V) == 0:
        return 0
    if len(V) == 1:
        a = abs(V.item())
        if a <= tol:
            return 0
        if abs(a - 1.0) <= tol:
            return 1
        raise RuntimeError("Invalid 1x1 V")
    s = np.linalg.svd(V, compute_uv=False)
    if len(V) > 2:
        return int((np.argmax(-np.diff(s)) + 1) % 2)                               # :445-446
    if np.allclose(s, [1.0, 0.0], rtol=0, atol=tol):
        return 1
    if abs(s[0] - s[1]) <= tol:
        return 0
    raise ValueError("Invalid 2x2 V")


# --------------------------------------------------------------------------------------------
# Schmidt modes of one bond (pfaffian.py:685-920)
# --------------------------------------------------------------------------------------------
@dataclass
class PfModes:
    nL: int
    nR: int
    e: np.ndarray
    vL: np.ndarray | None
    vR: np.ndarray | None
    pL: int | None
    pR: int | None

    def parity(self, which="T"):
        w = which[0].upper()
        if w == "L":
            return self.pL
        if w == "R":
            return self.pR
        return None if (self.pL is None or self.pR is None) else (self.pL + self.pR) % 2


def _diag_block(c, cutoff, deg_tol, diag_tol):
    """pfaffian.py:764-823: eigh of one diagonal Majorana block, clipped, 1/2-modes made real."""
    n = len(c) // 2
    if n == 0:
        return np.zeros(0), np.zeros((0, 0), c.dtype), 0, 0
    e, v = np.linalg.eigh(c)
    e = np.clip(e, 0.0, 1.0)                                                       # :791-796
    lo, hi = np.searchsorted(e, [0.5 - deg_tol, 0.5 + deg_tol])                    # :803
    kh = hi - n
    assert lo == n - kh, "1/2 eigenvalues asymmetrical in spectrum"
    if kh and np.iscomplexobj(v):                                                  # :808-816
        w = np.column_stack((v[:, lo:hi].real, v[:, lo:hi].imag))
        w, s, _ = np.linalg.svd(w)
        v[:, lo:hi] = w[:, : 2 * kh]
    lo, hi = np.searchsorted(e, [cutoff, 1 - cutoff])                              # :819
    ke = hi - n
    assert lo == n - ke, "Entangled modes asymmetrical in spectrum"
    return e, v, int(ke), int(kh)


def _restore_nambu(v, kh, side):
    """pfaffian.py:880-897: conjugate pairs for the 1/2 modes, Nambu completion, complex-fermion
    basis and the parity of the vacuum."""
    x = len(v) // 2
    if side == "L":
        v[:, x - kh: x] = (v[:, x - kh: x] + 1j * v[:, x: x + kh]) / 2 ** 0.5
        v[:, x:] = v[:, :x].conj()
    else:
        v[:, x: x + kh] = (-1j * v[:, x - kh: x] + v[:, x: x + kh]) / 2 ** 0.5
        v[:, x: x + kh] = v[:, x: x + kh][:, ::-1]
        v[:, :x] = v[:, x:].conj()
    v = vector_M2C(v)
    return v, vacuum_parity(v[1::2, :x])


def bond_modes(C, x, trunc, basis, which="LR", total_parity=None, diag_tol=1e-8) -> PfModes:
    trunc = Trunc.make(trunc)
    cutoff = trunc.svd_min ** 2
    deg_tol = trunc.degeneracy_tol
    if basis == "C":
        C = matrix_C2M(C)
    elif basis != "M":
        raise ValueError("`basis` must be 'M' or 'C'")
    C = regularise_nambu(C, "M", 1)                                                # :754
    L = len(C) // 2
    y = L - x
    which = which.upper()
    eL = vL = eR = vR = None
    if "L" in which:
        eL, vL, keL, khL = _diag_block(C[: 2 * x, : 2 * x], cutoff, deg_tol, diag_tol)
    if "R" in which:
        eR, vR, keR, khR = _diag_block(C[2 * x:, 2 * x:], cutoff, deg_tol, diag_tol)
    if eL is None:
        k, kh, e = keR, khR, eR[y - keR: y]
    elif eR is None:
        k, kh, e = keL, khL, eL[x - keL: x]
    else:
        assert keL == keR and khL == khR
        k, kh, e = keL, khL, eL[x - keL: x]
        CLR = C[: 2 * x, 2 * x:]
        vLE = vL[:, x - k: x - kh]                                                 # views, rotated in place
        vRE = vR[:, y + kh: y + k][:, ::-1]
        block_svd(CLR, vLE, vRE, eL[x - k: x - kh], deg_tol)                       # :855
        if kh:                                                                     # :860-865
            sl, sr = slice(x - kh, x + kh), slice(y - kh, y + kh)
            blk = vL[:, sl].real.T @ CLR.imag @ vR[:, sr].real
            U, _, Vh = np.linalg.svd(blk)
            vL[:, sl] = vL[:, sl] @ U
            vR[:, sr] = vR[:, sr] @ Vh.T
    if kh > 0:                                                                     # :868-874
        from scipy.stats import ortho_group
        O = ortho_group.rvs(2 * kh, random_state=1234)
        if vL is not None:
            vL[:, x - kh: x + kh] = vL[:, x - kh: x + kh] @ O
        if vR is not None:
            vR[:, y - kh: y + kh] = vR[:, y - kh: y + kh] @ O
    pL = pR = None
    if "L" in which:
        vL, pL = _restore_nambu(vL, kh, "L")
        if "R" not in which and total_parity is not None:
            pR = (total_parity + pL) % 2
    if "R" in which:
        vR, pR = _restore_nambu(vR, kh, "R")
        if "L" not in which and total_parity is not None:
            pL = (total_parity + pR) % 2
    if "L" in which and "R" in which and pL == 1:                                  # :915-916
        vR = -vR
    return PfModes(nL=x, nR=y, e=e, vL=vL, vR=vR, pL=pL, pR=pR)


# --------------------------------------------------------------------------------------------
# Schmidt vectors (pfaffian.py:986-1005, 1162-1214)
# --------------------------------------------------------------------------------------------
def parity_n_argsort(x):
    """Stable order by (parity, value) and the bunched slices of both keys (pfaffian.py:986-1005)."""
    x = np.asarray(x).ravel()
    idx = np.lexsort((np.arange(len(x)), x, x % 2))
    xs = x[idx]
    return idx, _bunch(xs), _bunch(xs % 2)


def _bunch(x):
    cuts = np.concatenate(([0], np.flatnonzero(x[1:] != x[:-1]) + 1, [len(x)]))
    return {int(x[cuts[i]]): slice(int(cuts[i]), int(cuts[i + 1])) for i in range(len(cuts) - 1)}


@dataclass
class PfVectors:
    modes: PfModes
    sets: np.ndarray             # (chi, k) bool in the order of ``e`` (left order)
    lam: np.ndarray
    idx_n: dict
    idx_parity: dict

    def side_sets(self, mode):
        return self.sets if mode[0].lower() == "l" else self.sets[:, ::-1]         # :955-957

    def charges(self, p_vac):
        """parity charge of every Schmidt vector = (excitation parity + vacuum parity) % 2
        (pfaffian.py:1485-1489)."""
        q = np.zeros(len(self.lam), dtype=np.int64)
        for par, slc in self.idx_parity.items():
            q[slc] = (par + p_vac) % 2
        return q


def bond_vectors(modes: PfModes, trunc) -> PfVectors:
    trunc = Trunc.make(trunc)
    a = np.log((1 - modes.e) / modes.e) / 2                                        # :925, :1189
    _, sets = lowest_sums(a, trunc)
    if len(sets) == 0:
        raise ValueError("No Schmidt vectors left after filtering by `trunc_par.sectors`!")
    idx, idx_n, idx_parity = parity_n_argsort(sets.sum(axis=1))
    sets = sets[idx]
    lam = np.where(sets, modes.e, 1 - modes.e).prod(axis=1) ** 0.5                 # :979
    return PfVectors(modes=modes, sets=sets, lam=lam, idx_n=idx_n, idx_parity=idx_parity)


def bond_vectors_from_C(C, x, trunc, basis, which="LR", total_parity=None) -> PfVectors:
    return bond_vectors(bond_modes(C, x, trunc, basis, which, total_parity), trunc)


# --------------------------------------------------------------------------------------------
# Pfaffians: Parlett-Reid with partial pivoting, batched over leading axes (replaces pfapack)
# --------------------------------------------------------------------------------------------
def pfaffian(A):
    """Pf of antisymmetric matrices ``A[..., m, m]`` (only the strict upper triangle is trusted,
    like pfapack's ``uplo="U"``)."""
    A = np.array(A, dtype=np.result_type(np.asarray(A).dtype, np.float64))
    m = A.shape[-1]
    batch = A.shape[:-2]
    if m == 0:
        return np.ones(batch, dtype=A.dtype)
    A = A.reshape(-1, m, m).copy()
    iu = np.triu_indices(m, 1)
    A[:, iu[1], iu[0]] = -A[:, iu[0], iu[1]]
    A[:, np.arange(m), np.arange(m)] = 0
    nb = len(A)
    if m % 2:
        return np.zeros(batch, dtype=A.dtype)
    pf = np.ones(nb, dtype=A.dtype)
    ar = np.arange(nb)
    for k in range(0, m - 1, 2):
        piv = k + 1 + np.argmax(np.abs(A[:, k, k + 1:]), axis=1)
        swap = piv != k + 1
        if np.any(swap):
            rows = A[ar, piv].copy()
            A[ar, piv] = A[:, k + 1]
            A[:, k + 1] = rows
            cols = A[ar, :, piv].copy()
            A[ar, :, piv] = A[:, :, k + 1]
            A[:, :, k + 1] = cols
            pf = np.where(swap, -pf, pf)
        p = A[:, k, k + 1]
        pf = pf * p
        if k + 2 < m:
            with np.errstate(divide="ignore", invalid="ignore"):
                tau = np.where(p[:, None] != 0, A[:, k, k + 2:] / p[:, None], 0)
            col = A[:, k + 2:, k + 1]
            upd = tau[:, :, None] * col[:, None, :]
            A[:, k + 2:, k + 2:] += upd - np.transpose(upd, (0, 2, 1))
    return pf.reshape(batch)


# --------------------------------------------------------------------------------------------
# MPS tensors (pfaffian.py:1258-1479, 1578-1778)
# --------------------------------------------------------------------------------------------
def pfaffian_matrix(V1, V2, sets1, sets2, mode):
    """pfaffian.py:1258-1410: Onishi norm and the antisymmetric contraction matrix
    ``N = [[BB, BA], [-BA^T, AA]]`` over the active ket (b^+) and bra (a) modes."""
    L = V1.shape[0] // 2
    Vr = V1.conj().T @ V2
    s = np.linalg.svd(Vr[:L, :L], compute_uv=False)
    norm = s.prod() ** 0.5

    def prune(sets, reverse):
        idx = np.flatnonzero(np.any(sets, axis=0))
        if reverse:
            idx = idx[::-1]
        return sets[:, idx], idx

    act1, act2 = sets1.shape[1], sets2.shape[1]
    sets1, idx1 = prune(sets1, False)
    sets2, idx2 = prune(sets2, True)
    if mode == "left":
        idx1 = idx1 + (L - act1)
        idx2 = idx2 + (L - act2)
    Ux = np.linalg.inv(Vr[L:, L:])
    AA = Vr[idx1, L:] @ Ux[:, idx1]
    BA = Ux[np.ix_(idx2, idx1)]
    BB = Ux[idx2] @ Vr[L:, idx2]
    AA = (AA - AA.T) / 2
    BB = (BB - BB.T) / 2
    N = np.block([[BB, BA], [-BA.T, AA]])
    n1 = np.concatenate((np.zeros((len(sets1), sets2.shape[1]), bool), sets1), axis=1)
    n2 = np.concatenate((sets2, np.zeros((len(sets2), sets1.shape[1]), bool)), axis=1)
    return norm, N, n1, n2


def tensor_block(N, sets1, sets2):
    """pfaffian.py:1429-1479: Pf(N[idx, idx]) with idx = ket excitations ++ bra excitations."""
    n1 = int(sets1[0].sum())
    n2 = int(sets2[0].sum())
    r1 = np.nonzero(sets1)[1].reshape(len(sets1), n1)
    r2 = np.nonzero(sets2)[1].reshape(len(sets2), n2)
    idx = np.concatenate((np.broadcast_to(r2[None], (len(r1), len(r2), n2)),
                          np.broadcast_to(r1[:, None], (len(r1), len(r2), n1))), axis=-1)
    if n1 + n2 == 0:
        return np.ones((len(r1), len(r2)), dtype=N.dtype)
    return pfaffian(N[idx[..., :, None], idx[..., None, :]])


@dataclass
class PfTensorData:
    mode: str
    norm: float
    N: np.ndarray
    qtotal: int
    sets_bra: np.ndarray        # sorted by (parity, n)
    sets_ket: np.ndarray
    idx_n_bra: dict
    idx_n_ket: dict
    leg_idx_bra: np.ndarray     # sorted row -> pipe row (p * chi_bra + alpha), or alpha without p
    chi_bra: int
    chi_ket: int
    physical: bool
    q_bra: np.ndarray           # parity charge of every bra Schmidt vector
    q_ket: np.ndarray


def tensor_data(bra: PfVectors, ket: PfVectors, mode: str) -> PfTensorData:
    """pfaffian.py:1578-1748."""
    mode = mode.lower()
    side = "L" if mode == "left" else "R"
    mb, mk = bra.modes, ket.modes
    v_bra = (mb.vL if mode == "left" else mb.vR)
    v_ket = (mk.vL if mode == "left" else mk.vR)
    sets_bra = bra.side_sets(mode)
    p_bra, p_ket = mb.pL, mk.pL
    if p_bra is None or p_ket is None:
        p_bra, p_ket, qtotal = mb.pR, mk.pR, 0
    elif mode == "right":
        qtotal = (mb.parity() + mk.parity()) % 2                                   # :1641
    else:
        qtotal = 0
    chi_bra = len(sets_bra)
    physical = False
    if len(v_bra) + 2 == len(v_ket):                                               # :1650-1694
        physical = True
        n = len(v_bra) // 2
        zc, zr = np.zeros((2 * n, 1)), np.zeros((1, n))
        off, on = np.zeros((chi_bra, 1), bool), np.ones((chi_bra, 1), bool)
        if mode == "left":
            u = -1 if mb.parity(side) % 2 == 1 else 1
            v_bra = np.block([[v_bra[:, :n], zc, v_bra[:, n:], zc],
                              [zr, u, zr, 0.0],
                              [zr, 0.0, zr, u]])
            sets_bra = np.block([[sets_bra, off], [sets_bra, on]])
        else:
            v_bra = np.block([[1, zr, 0, zr],
                              [0, zr, 1, zr],
                              [zc, v_bra[:, :n], zc, v_bra[:, n:]]])
            sets_bra = np.block([[off, sets_bra], [on, sets_bra]])
    elif len(v_bra) == len(v_ket):
        v_bra, sets_bra = v_bra.copy(), sets_bra.copy()
    else:
        raise ValueError("bra and ket sizes do not match")
    if mb.parity(side) % 2 != mk.parity(side) % 2:                                 # :1708-1719
        n = len(v_bra) // 2
        v_bra = v_bra.astype(complex)
        if mode == "left":
            v_bra[:, [n - 1, 2 * n - 1]] = v_bra[:, [2 * n - 1, n - 1]]
            sets_bra[:, -1] = ~sets_bra[:, -1]
        else:
            v_bra *= -1
            v_bra[:, [0, n]] = -v_bra[:, [n, 0]]
            sets_bra[:, 0] = ~sets_bra[:, 0]
    norm, N, sets_bra, sets_ket = pfaffian_matrix(v_bra, v_ket, sets_bra, ket.side_sets(mode), mode)
    leg_idx, idx_n_bra, _ = parity_n_argsort(sets_bra.sum(axis=1))                 # :1732
    return PfTensorData(mode=mode, norm=norm, N=N, qtotal=qtotal, sets_bra=sets_bra[leg_idx],
                        sets_ket=sets_ket, idx_n_bra=idx_n_bra, idx_n_ket=ket.idx_n,
                        leg_idx_bra=leg_idx, chi_bra=chi_bra, chi_ket=len(sets_ket), physical=physical,
                        q_bra=bra.charges(p_bra), q_ket=ket.charges(p_ket))


def dense_tensor(td: PfTensorData):
    """pfaffian.py:1750-1778 densely: ``T[p, alpha, beta]`` (or ``T[alpha, beta]``)."""
    M = np.zeros((len(td.sets_bra), td.chi_ket), dtype=complex)
    for nb, sb in td.idx_n_bra.items():
        for nk, sk in td.idx_n_ket.items():
            if (nb + nk) % 2 == 1:
                continue
            M[td.leg_idx_bra[sb], sk] = td.norm * tensor_block(td.N, td.sets_bra[sb], td.sets_ket[sk])
    if not td.physical:
        return M
    return M.reshape(2, td.chi_bra, td.chi_ket)


# --------------------------------------------------------------------------------------------
# chain driver (pfaffian.py:1785-1921)
# --------------------------------------------------------------------------------------------
def C_to_MPS(C, trunc, basis, ortho_center=None) -> DenseMPS:
    trunc = Trunc.make(trunc)
    L = len(C) // 2
    oc = ortho_center or L // 2
    tensors, lams, charges = [None] * L, [None] * (L + 1), [None] * (L + 1)

    def unit(v):
        return v / np.linalg.norm(v)

    centre = bond_vectors_from_C(C, oc, trunc, basis)
    lams[oc], charges[oc] = unit(centre.lam), centre.charges(centre.modes.pL)
    parity = centre.modes.parity()
    prev = centre
    for i in range(oc, L):                                                         # :1856-1883
        new = bond_vectors_from_C(C, i + 1, trunc, basis, "R", parity)
        lams[i + 1], charges[i + 1] = unit(new.lam), new.charges(new.modes.pL)
        T = dense_tensor(tensor_data(new, prev, "right"))
        tensors[i] = np.transpose(T, (2, 0, 1))
        prev = new
    prev = centre
    for i in reversed(range(oc)):                                                  # :1887-1914
        new = bond_vectors_from_C(C, i, trunc, basis, "L", parity)
        lams[i], charges[i] = unit(new.lam), new.charges(new.modes.pL)
        T = dense_tensor(tensor_data(new, prev, "left"))
        tensors[i] = np.transpose(T, (1, 0, 2))
        prev = new
    return DenseMPS(tensors=tensors, lams=lams, charges=charges, form=["A"] * oc + ["B"] * (L - oc),
                    ortho_center=oc)


# --------------------------------------------------------------------------------------------
# known answers
# --------------------------------------------------------------------------------------------
def state_correlators(psi):
    """<c_i^+ c_j> and <c_i c_j> of a dense fermion state psi[n_0..n_{L-1}] (Jordan-Wigner, site 0
    leftmost / first created).  Returns (G, F) with G[i, j] = <c_i^+ c_j>, F[i, j] = <c_i c_j>."""
    L = psi.ndim
    psi = psi / np.linalg.norm(psi)

    def apply(op, i, s):
        """c_i (op=0) or c_i^+ (op=1) on state s."""
        out = np.zeros_like(s)
        src = [slice(None)] * L
        dst = [slice(None)] * L
        src[i], dst[i] = (1, 0) if op == 0 else (0, 1)
        out[tuple(dst)] = s[tuple(src)]
        sign = np.ones((2,) * L)
        for j in range(i):
            idx = [slice(None)] * L
            idx[j] = 1
            sign[tuple(idx)] *= -1
        return out * sign

    G = np.zeros((L, L), dtype=complex)
    F = np.zeros((L, L), dtype=complex)
    for i in range(L):
        for j in range(L):
            G[i, j] = np.vdot(psi, apply(1, i, apply(0, j, psi)))
            F[i, j] = np.vdot(psi, apply(0, i, apply(0, j, psi)))
    return G, F


# --------------------------------------------------------------------------------------------
# iMPS driver (pfaffian.py:1924-2091)
# --------------------------------------------------------------------------------------------
@dataclass
class PfDenseIMPS:
    tensors: list          # T[vL, p, vR], right-canonical ("B" form)
    lams: list             # cell + 1 normalised Schmidt vectors
    charges: list          # cell + 1 parity arrays (parity left of the bond)
    errors: tuple          # (left_unitary, left_schmidt, 0.0, 0.0)
    qtotal: list


def C_to_iMPS(C_short, C_long, trunc, sites_per_cell, cut, basis) -> PfDenseIMPS:
    """pfaffian.py:1924-2091 with dense tensors: Schmidt vectors of both chains at ``cut`` (:1998-2005), right
    tensors of the additional unit cell of the long chain with the short chain's right environment closing the cell
    (:2009-2048), gauge fixing of the first tensor by the overlap of the two left Schmidt bases (:2051-2063)."""
    from slater_oracle import basis_rotation
    trunc = Trunc.make(trunc)
    assert len(C_short) // 2 + sites_per_cell == len(C_long) // 2

    def unit(v):
        return v / np.linalg.norm(v)

    short = bond_vectors_from_C(C_short, cut, trunc, basis)
    long_ = bond_vectors_from_C(C_long, cut, trunc, basis)
    lams, tensors, qtot = [unit(short.lam)], [], []
    charges = [long_.charges(long_.modes.pL)]
    prev = long_
    for i in range(sites_per_cell):
        if i == sites_per_cell - 1:
            new = short
            lams.append(lams[0])
        else:
            new = bond_vectors_from_C(C_long, cut + i + 1, trunc, basis, "R", long_.modes.parity())
            lams.append(unit(new.lam))
        td = tensor_data(new, prev, "right")
        tensors.append(np.transpose(dense_tensor(td), (2, 0, 1)))
        charges.append(td.q_bra)
        qtot.append(td.qtotal)
        prev = new
    td = tensor_data(short, long_, "left")                                        # no physical leg
    R, uerr, serr = basis_rotation(dense_tensor(td), td.q_bra, td.q_ket, short.lam, long_.lam)
    tensors[0] = np.tensordot(R, tensors[0], axes=(1, 0))
    charges[0] = td.q_bra
    return PfDenseIMPS(tensors=tensors, lams=lams, charges=charges, errors=(uerr, serr, 0.0, 0.0), qtotal=qtot)
