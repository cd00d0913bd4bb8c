"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Slater-determinant -> MPS hot path.

A NumPy restatement of the reference algorithm (temfpy/temfpy, read-only checkout at
/root/reference; every function cites the reference ``file:line`` it follows).  It is the checker
for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import it.  Nothing under ``temfpy_b200/`` does.

Parity pins (see DESIGN.md "Oracle"):
  * the reference ships no tests or golden vectors ("parity unpinned" by the reference itself);
  * this restatement is therefore pinned (a) against the reference's own code imported in the build
    container through ``oracle/ref_shim.py`` -- fixtures under ``tests/golden/`` made by
    ``oracle/make_golden.py`` -- and (b) against exact known answers: Slater amplitudes
    ``det(Phi[occ, :])`` for small chains (``exact_slater_state``).

TeNPy is not installed in this image, so the reference's ``to_npc_array``/``MPS(...)`` packing
(slater.py:1106-1143, 1348-1351) is restated densely: every site tensor is returned as an ndarray
``T[vL, p, vR]`` together with the integer charge (fermion number to the left) of every virtual
index, which is exactly the information TeNPy's block structure carries.
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass, field

import numpy as np

DEFAULT_SVD_MIN = 1e-6      # schmidt_utils.py:14
DEFAULT_DEG_TOL = 1e-12     # schmidt_utils.py:15


# --------------------------------------------------------------------------------------------
# truncation rule  (schmidt_utils.py:18-208)
# --------------------------------------------------------------------------------------------
class Trunc:
    """Stopping/truncation rule; mirrors ``StoppingCondition`` (schmidt_utils.py:18-185)."""

    def __init__(self, sectors=None, chi_max=None, svd_min=None, degeneracy_tol=None):
        self.chi_max = chi_max
        self.svd_min = DEFAULT_SVD_MIN if svd_min is None else svd_min            # :58-60
        self.degeneracy_tol = DEFAULT_DEG_TOL if degeneracy_tol is None else degeneracy_tol  # :63-65
        self.sectors = sectors
        if sectors is None:                                                        # :68-78
            self.is_sector = lambda q: True
        elif isinstance(sectors, (int, np.integer, float)):
            self.is_sector = lambda q: q == sectors
        elif callable(sectors):
            self.is_sector = sectors
        else:
            allowed = list(sectors)
            self.is_sector = lambda q: q in allowed
        assert self.chi_max is None or self.chi_max > 0
        assert 0 < self.svd_min < 1
        assert self.degeneracy_tol > 0
        self.max_logval = -np.log(self.svd_min) + self.degeneracy_tol              # :96

    @classmethod
    def make(cls, tp):
        if isinstance(tp, cls):
            return tp
        if isinstance(tp, dict):
            return cls(**tp)
        # duck-typed StoppingCondition-like object
        return cls(sectors=getattr(tp, "sectors", None), chi_max=tp.chi_max, svd_min=tp.svd_min,
                   degeneracy_tol=tp.degeneracy_tol)

    def more_needed(self, logvals) -> bool:
        """schmidt_utils.py:99-138 -- chi_max+1 look-ahead and dynamic-range stop."""
        if self.chi_max is not None and len(logvals) > self.chi_max:
            return False
        if logvals[-1] - logvals[0] > self.max_logval:
            return False
        return True

    def truncate(self, logvals) -> int:
        """schmidt_utils.py:140-185 -- last admissible cut position + 1."""
        lv = np.asarray(logvals, dtype=float)
        ok = np.ones(lv.size, dtype=bool)
        if self.chi_max is not None:
            ok[self.chi_max:] = False
        ok &= (lv - lv[0]) < -np.log(self.svd_min)
        gap_ok = np.ones(lv.size, dtype=bool)
        gap_ok[:-1] = (lv[1:] - lv[:-1]) > self.degeneracy_tol
        ok &= gap_ok
        return int(np.flatnonzero(ok)[-1]) + 1


def lowest_sums(a, trunc: Trunc, filled_left=None, filled_right=None):
    """Best-first enumeration of the subsets of ``a`` with the lowest sums.

    Follows schmidt_utils.py:211-324: start from the set of negative entries (:274-275), children
    are "flip the next larger |a|" and "move the last flip to the next larger |a|" (:304-315),
    ties broken by push order (:290, :308, :314), loop while the stopping rule asks for more
    (:297), sector filter on the left (or right) particle number (:257-266, :277, :300), final
    ``truncate`` (:321).
    """
    a = np.asarray(a, dtype=float)
    k = a.size

    def charge(s):
        n = int(np.count_nonzero(s))
        if filled_left is not None:
            return filled_left + n
        if filled_right is not None:
            return filled_right + k - n
        return n

    if k == 0:                                                                     # :268-271
        keep = int(bool(trunc.is_sector(charge(np.zeros(0, bool)))))
        return np.zeros(keep), np.zeros((keep, 0), bool)

    neg = a < 0
    base = np.sum(a[neg])                      # numpy pairwise sum, as the reference (:274)
    sums, sets = [], []
    if trunc.is_sector(charge(neg)):
        sums.append(base)
        sets.append(neg.copy())
    mag = np.abs(a)
    order = np.argsort(mag)                                                        # :286
    first = neg.copy()
    first[order[0]] ^= True
    heap = [(base + mag[order[0]], 0, 0, first)]
    seq = 0
    # NB the reference evaluates ``trunc_par(sums)`` on a possibly empty list when the lowest
    # set is filtered out by ``sectors`` -> IndexError.  We guard it instead.
    while heap and (len(sums) == 0 or trunc.more_needed(sums)):
        s, _, i, cur = heapq.heappop(heap)
        if trunc.is_sector(charge(cur)):
            sums.append(s)
            sets.append(cur)
        if i < k - 1:
            c1 = cur.copy()
            c1[order[i + 1]] ^= True
            s = s + mag[order[i + 1]]
            seq += 1
            heapq.heappush(heap, (s, seq, i + 1, c1))
            c2 = c1.copy()
            c2[order[i]] ^= True
            s = s - mag[order[i]]
            seq += 1
            heapq.heappush(heap, (s, seq, i + 1, c2))
    if not sums:
        return np.zeros(0), np.zeros((0, k), bool)
    sums = np.asarray(sums)
    sets = np.asarray(sets)
    cut = trunc.truncate(sums)
    return sums[:cut], sets[:cut]


# --------------------------------------------------------------------------------------------
# correlation matrices (slater.py:1150-1213)
# --------------------------------------------------------------------------------------------
def correlation_matrix(H, N=None):
    """slater.py:1150-1180: C = Phi Phi^dagger of the N lowest (default: negative) levels."""
    w, v = np.linalg.eigh(H)
    if N is None:
        occ = w < 0
        v = v[:, occ]
        N = int(occ.sum())
    else:
        v = v[:, :N]
    C = v @ v.conj().T
    if np.iscomplexobj(C) and np.allclose(C.imag, 0.0, rtol=0, atol=1e-14):
        C = C.real
    return C, N


def spinful_correlation_matrix(C, ph=True):
    """slater.py:1183-1213: up spins on even, down spins on odd sites; PH: 1-C for down."""
    n = len(C)
    C2 = np.zeros((2 * n, 2 * n), dtype=C.dtype)
    C2[::2, ::2] = C
    C2[1::2, 1::2] = (np.eye(n) - C) if ph else C
    return C2


# --------------------------------------------------------------------------------------------
# Schmidt modes of one bond (slater.py:270-423, utils.py:19-96)
# --------------------------------------------------------------------------------------------
@dataclass
class Modes:
    e: np.ndarray
    vL: np.ndarray | None
    vR: np.ndarray | None
    ixL: dict | None          # name -> (start, stop)
    ixR: dict | None
    nL: int
    nR: int
    n_fermion: int

    def n_filled(self, side):                                                      # :145-174
        k = self.e.size
        if side == "L":
            if self.ixL is not None:
                return self.ixL["filled"][1] - self.ixL["filled"][0]
            return self.n_fermion - k - (self.ixR["filled"][1] - self.ixR["filled"][0])
        if self.ixR is not None:
            return self.ixR["filled"][1] - self.ixR["filled"][0]
        return self.n_fermion - k - (self.ixL["filled"][1] - self.ixL["filled"][0])


def _split_block(c, side, cutoff):
    """eigh of a diagonal block + split into filled / entangled / empty (slater.py:324-375)."""
    n = len(c)
    if n == 0:
        z = (0, 0)
        return np.zeros(0), np.zeros((0, 0), c.dtype), dict(filled=z, entangled=z, empty=z), 0
    w, v = np.linalg.eigh(c)
    lo, hi = np.searchsorted(w, [cutoff, 1 - cutoff])                              # :350
    n_empty, k, n_fill = lo, hi - lo, n - hi
    perm = np.arange(n)
    if side == "L":        # filled, entangled (decreasing), empty                 # :355-361
        perm = perm[::-1]
        ix = dict(filled=(0, n_fill), entangled=(n_fill, n_fill + k), empty=(n_fill + k, n))
    else:                  # empty, entangled (decreasing), filled                 # :362-368
        perm[lo:hi] = perm[lo:hi][::-1]
        ix = dict(empty=(0, lo), entangled=(lo, hi), filled=(hi, n))
    w = w[perm]
    v = v[:, perm]
    return w[ix["entangled"][0]:ix["entangled"][1]], v, ix, int(k)


def block_svd(CLR, vL, vR, e, deg_tol):
    """utils.py:19-96: SVD inside groups of (nearly) degenerate ``e``; rotates vL, vR in place."""
    k = e.size
    if k == 0:
        return
    breaks = np.flatnonzero(np.abs(np.diff(e)) > deg_tol) + 1
    starts = np.concatenate(([0], breaks))
    stops = np.concatenate((breaks, [k]))
    for a, b in zip(starts, stops):
        blockL = vL[:, a:b]
        blockR = vR[:, a:b]
        s = blockL.conj().T @ CLR @ blockR
        U, _, Vh = np.linalg.svd(s)
        vL[:, a:b] = blockL @ U
        vR[:, a:b] = blockR @ Vh.conj().T


def bond_modes(C, x, trunc: Trunc, which="LR") -> Modes:
    """slater.py:270-423."""
    cutoff = trunc.svd_min ** 2                                                    # :318
    which = which.upper()
    L = len(C)
    eL = vL = ixL = eR = vR = ixR = None
    if "L" in which:
        eL, vL, ixL, kL = _split_block(C[:x, :x], "L", cutoff)
    if "R" in which:
        eR, vR, ixR, kR = _split_block(C[x:, x:], "R", cutoff)
    if eL is None:
        e = 1.0 - eR[::-1]                                                         # :386
    elif eR is None:
        e = eL
    else:
        assert kL == kR, "number of entangled modes differs between the two sides"  # :394
        e = eL
        a, b = ixL["entangled"]
        c, d = ixR["entangled"]
        vLE = vL[:, a:b]                     # views: block_svd rotates in place
        vRE_rev = vR[:, c:d][:, ::-1]
        block_svd(C[:x, x:], vLE, vRE_rev, e, trunc.degeneracy_tol)                # :407
        vR[:, c:d][:, 1::2] *= -1                                                  # :410
    n_fermion = int(np.round(np.trace(C).real))                                    # :414
    return Modes(e=e, vL=vL, vR=vR, ixL=ixL, ixR=ixR, nL=x, nR=L - x, n_fermion=n_fermion)


# --------------------------------------------------------------------------------------------
# Schmidt vectors of one bond (slater.py:430-489, 633-700)
# --------------------------------------------------------------------------------------------
@dataclass
class Vectors:
    modes: Modes
    sets: np.ndarray              # (chi, k) bool: entangled mode occupied on the LEFT
    left_sets: np.ndarray | None  # (chi, nL)
    right_sets: np.ndarray | None  # (chi, nR)
    lam: np.ndarray               # (chi,)  un-normalised Schmidt values
    n_left: np.ndarray            # (chi,) fermion number to the left
    idx_L: dict = field(default_factory=dict)   # charge -> (start, stop), ascending charge


def bond_vectors(modes: Modes, trunc: Trunc) -> Vectors:
    """slater.py:633-700 (+ embed_subsets :430-470, schmidt_values :472-489)."""
    e = modes.e
    ratio = np.log((1 - e) / e)                                                    # :428
    _, sets = lowest_sums(ratio / 2, trunc, filled_left=modes.n_filled("L"),
                          filled_right=modes.n_filled("R"))                        # :662-667
    if len(sets) == 0:
        raise ValueError("No Schmidt vectors left after filtering by `trunc_par.sectors`!")
    nL = modes.n_filled("L") + sets.sum(axis=1)                                    # :673
    order = np.argsort(nL, kind="stable")                                          # :676
    nL = nL[order]
    sets = sets[order]
    q, first = np.unique(nL, return_index=True)                                    # :681
    bounds = np.concatenate((first, [len(sets)]))
    idx_L = {int(q[i]): (int(bounds[i]), int(bounds[i + 1])) for i in range(len(q))}
    left_sets = right_sets = None
    if modes.vL is not None:                                                       # :456-461
        left_sets = np.zeros((len(sets), modes.nL), bool)
        a, b = modes.ixL["entangled"]
        left_sets[:, a:b] = sets
        a, b = modes.ixL["filled"]
        left_sets[:, a:b] = True
    if modes.vR is not None:                                                       # :463-468
        right_sets = np.zeros((len(sets), modes.nR), bool)
        a, b = modes.ixR["entangled"]
        right_sets[:, a:b] = ~sets[:, ::-1]
        a, b = modes.ixR["filled"]
        right_sets[:, a:b] = True
    lam = np.where(sets, e, 1 - e).prod(axis=1) ** 0.5                             # :489
    return Vectors(modes=modes, sets=sets, left_sets=left_sets, right_sets=right_sets, lam=lam,
                   n_left=nL, idx_L=idx_L)


def bond_vectors_from_C(C, x, trunc, which="LR") -> Vectors:
    """slater.py:702-755."""
    trunc = Trunc.make(trunc)
    return bond_vectors(bond_modes(C, x, trunc, which), trunc)


# --------------------------------------------------------------------------------------------
# site tensors (slater.py:760-1143)
# --------------------------------------------------------------------------------------------
def select_orbitals(sets, V, mode):
    """slater.py:760-825: keep always+sometimes occupied orbitals, with reordering signs."""
    always = np.flatnonzero(sets.all(axis=0))
    sometimes = np.flatnonzero(sets.any(axis=0) & ~sets.all(axis=0))
    k = always.size
    n_before = np.searchsorted(always, sometimes)     # always-orbitals left of each sometimes one
    if mode == "left":
        cols = np.concatenate((always, sometimes))
        sign = np.concatenate((np.ones(k), (-1.0) ** (k - n_before)))              # :813
    else:
        cols = np.concatenate((sometimes, always))
        sign = np.concatenate(((-1.0) ** n_before, np.ones(k)))                    # :820
    return sets[:, cols], V[:, cols] * sign, k


@dataclass
class TensorData:
    mode: str
    physical: bool
    det_always: complex
    S: np.ndarray                 # "sometimes matrix"
    sets_bra: np.ndarray          # (2 chi_b or chi_b, s_b) bool, rows in the reference's bra order
    sets_ket: np.ndarray          # (chi_k, s_k) bool
    bra_rows: np.ndarray          # for every bra row: (p, alpha)  [p = -1 without physical leg]
    q_bra: np.ndarray             # pipe charge of every bra row (charge to the left)
    q_ket: np.ndarray             # charge of every ket row
    qtotal: int
    chi_bra: int
    chi_ket: int


def tensor_data(bra: Vectors, ket: Vectors, mode: str) -> TensorData:
    """slater.py:975-1104 (+ the row bookkeeping TeNPy's LegPipe does in :1113-1119)."""
    side = "L" if mode == "left" else "R"
    v_bra = bra.modes.vL if side == "L" else bra.modes.vR
    v_ket = ket.modes.vL if side == "L" else ket.modes.vR
    sb = bra.left_sets if side == "L" else bra.right_sets
    sk = ket.left_sets if side == "L" else ket.right_sets
    chi_b, n_b = sb.shape
    alpha = np.arange(chi_b)
    if n_b == sk.shape[1]:
        physical = False
        rows_p = np.full(chi_b, -1)
        rows_a = alpha
        q_bra = bra.n_left.copy()
    elif n_b + 1 == sk.shape[1]:
        physical = True
        ext = np.zeros((n_b + 1, n_b + 1), dtype=v_bra.dtype)
        empty = np.zeros((chi_b, 1), bool)
        full = np.ones((chi_b, 1), bool)
        if mode == "left":       # physical orbital appended at the end            # :1030-1040
            ext[:n_b, :n_b] = v_bra
            ext[n_b, n_b] = 1
            sb = np.block([[sb, empty], [sb, full]])
        else:                    # physical orbital prepended                       # :1041-1051
            ext[0, 0] = 1
            ext[1:, 1:] = v_bra
            sb = np.block([[empty, sb], [full, sb]])
        v_bra = ext
        rows_p = np.repeat([0, 1], chi_b)
        rows_a = np.tile(alpha, 2)
        occ = sb.sum(axis=1)
        order = np.argsort(occ if mode == "left" else -occ, kind="stable")          # :1053-1058
        sb = sb[order]
        rows_p = rows_p[order]
        rows_a = rows_a[order]
        # charge to the left of the combined (site + bond) leg
        q_bra = bra.n_left[rows_a] + rows_p if mode == "left" else bra.n_left[rows_a] - rows_p
    else:
        raise ValueError("bra/ket sizes do not match")
    sb, v_bra, k_b = select_orbitals(sb, v_bra, mode)                              # :1066
    sk, v_ket, k_k = select_orbitals(sk, v_ket, mode)                              # :1067
    k = min(k_b, k_k)                                                              # :1069
    O = v_bra.conj().T @ v_ket                                                     # :1071
    if k == 0:
        det_always, S = 1.0, O
    elif mode == "left":                                                           # :1077-1082
        a = O[:k, :k]
        det_always = np.linalg.det(a)
        S = O[k:, k:] - O[k:, :k] @ np.linalg.inv(a) @ O[:k, k:]
        sb, sk = sb[:, k:], sk[:, k:]
    else:                                                                          # :1083-1090
        d = O[-k:, -k:]
        det_always = np.linalg.det(d)
        S = O[:-k, :-k] - O[:-k, -k:] @ np.linalg.inv(d) @ O[-k:, :-k]
        sb, sk = sb[:, :-k], sk[:, :-k]
    qtotal = 0 if mode == "left" else ket.modes.n_fermion - bra.modes.n_fermion    # :1092
    return TensorData(mode=mode, physical=physical, det_always=det_always, S=S, sets_bra=sb,
                      sets_ket=sk, bra_rows=np.stack([rows_p, rows_a], axis=1), q_bra=q_bra,
                      q_ket=ket.n_left.copy(), qtotal=qtotal, chi_bra=chi_b, chi_ket=len(sk))


def tensor_block(S, sets_bra, sets_ket):
    """slater.py:828-869: all minors det(S[rows(alpha)][:, cols(beta)]) of one charge block."""
    nb = sets_bra.sum(axis=1)
    nk = sets_ket.sum(axis=1)
    assert np.all(nb == nb[0]) and np.all(nk == nk[0]) and nb[0] == nk[0]
    n = int(nb[0])
    rows = np.nonzero(sets_bra)[1].reshape(len(sets_bra), n)
    cols = np.nonzero(sets_ket)[1].reshape(len(sets_ket), n)
    sub = S[rows[:, None, :, None], cols[None, :, None, :]]
    return np.linalg.det(sub)


def dense_tensor(td: TensorData):
    """Dense equivalent of ``to_npc_array`` (slater.py:1106-1143).

    Returns ``T[p, alpha, beta]`` (or ``T[alpha, beta]`` without physical leg); entry non-zero only
    where ``q_bra == q_ket + qtotal * qconj0`` (:1134), qconj0 = +1 (left) / -1 (right) (:1111).
    """
    qc = 1 if td.mode == "left" else -1
    dtype = np.result_type(td.S.dtype, np.asarray(td.det_always).dtype)
    M = np.zeros((len(td.sets_bra), td.chi_ket), dtype=dtype)
    for q in np.unique(td.q_ket):
        kr = np.flatnonzero(td.q_ket == q)
        br = np.flatnonzero(td.q_bra == q + td.qtotal * qc)
        if br.size == 0:
            continue
        M[np.ix_(br, kr)] = td.det_always * tensor_block(td.S, td.sets_bra[br], td.sets_ket[kr])
    if not td.physical:
        out = np.zeros((td.chi_bra, td.chi_ket), dtype=dtype)
        out[td.bra_rows[:, 1]] = M
        return out
    T = np.zeros((2, td.chi_bra, td.chi_ket), dtype=dtype)
    T[td.bra_rows[:, 0], td.bra_rows[:, 1]] = M
    return T


# --------------------------------------------------------------------------------------------
# chain driver (slater.py:1216-1353) with a neutral MPS container
# --------------------------------------------------------------------------------------------
@dataclass
class DenseMPS:
    tensors: list          # T[vL, p, vR]
    lams: list             # L+1 normalised Schmidt-value vectors
    charges: list          # L+1 int arrays: fermion number left of the bond for every index
    form: list             # "A"/"B" per site
    ortho_center: int

    @property
    def L(self):
        return len(self.tensors)

    @property
    def chi(self):
        return [len(l) for l in self.lams]


def C_to_MPS(C, trunc, ortho_center=None, spinful=None) -> DenseMPS:
    """slater.py:1216-1353 with dense tensors in (vL, p, vR) layout."""
    trunc = Trunc.make(trunc)
    if spinful == "simple":
        C = spinful_correlation_matrix(C, False)
    elif spinful == "PH":
        C = spinful_correlation_matrix(C, True)
    elif spinful is not None:
        raise ValueError("`spinful` must be 'simple', 'PH', or `None`")
    L = len(C)
    oc = ortho_center or L // 2                                                    # :1291
    tensors = [None] * L
    lams = [None] * (L + 1)
    charges = [None] * (L + 1)

    def norm(v):
        return v / np.linalg.norm(v)                                               # utils.py:99-103

    centre = bond_vectors_from_C(C, oc, trunc, "LR")
    lams[oc], charges[oc] = norm(centre.lam), centre.n_left
    prev = centre
    for i in range(oc, L):                                                         # :1301-1321
        new = bond_vectors_from_C(C, i + 1, trunc, "R")
        lams[i + 1], charges[i + 1] = norm(new.lam), new.n_left
        T = dense_tensor(tensor_data(new, prev, "right"))    # T[p, alpha(bond i+1), beta(bond i)]
        tensors[i] = np.transpose(T, (2, 0, 1))
        prev = new
    prev = centre
    for i in reversed(range(oc)):                                                  # :1326-1346
        new = bond_vectors_from_C(C, i, trunc, "L")
        lams[i], charges[i] = norm(new.lam), new.n_left
        T = dense_tensor(tensor_data(new, prev, "left"))     # T[p, alpha(bond i), beta(bond i+1)]
        tensors[i] = np.transpose(T, (1, 0, 2))
        prev = new
    form = ["A"] * oc + ["B"] * (L - oc)                                           # :1348
    return DenseMPS(tensors=tensors, lams=lams, charges=charges, form=form, ortho_center=oc)


# --------------------------------------------------------------------------------------------
# known answers and comparison helpers
# --------------------------------------------------------------------------------------------
def mps_to_state(mps: DenseMPS):
    """Contracts A...A diag(lam_oc) B...B into amplitudes psi[n_0, ..., n_{L-1}] (small L only)."""
    L, oc = mps.L, mps.ortho_center
    psi = np.ones((1, 1), dtype=complex)
    for i in range(L):
        if i == oc:
            psi = psi * mps.lams[oc][None, :]
        T = mps.tensors[i]
        psi = np.tensordot(psi, T, axes=(1, 0)).reshape(-1, T.shape[2])
    if oc == L:
        psi = psi * mps.lams[oc][None, :]
    return psi.reshape((2,) * L)


def exact_slater_state(Phi):
    """Amplitudes det(Phi[occ, :]) of the Slater determinant filling the columns of ``Phi``;
    basis state = (c_0^+)^{n_0} (c_1^+)^{n_1} ... |0>, site 0 is the most significant bit."""
    L, N = Phi.shape
    psi = np.zeros((2,) * L, dtype=complex)
    for conf in range(2 ** L):
        bits = [(conf >> (L - 1 - i)) & 1 for i in range(L)]
        if sum(bits) != N:
            continue
        occ = [i for i in range(L) if bits[i]]
        psi[tuple(bits)] = np.linalg.det(Phi[occ, :])
    return psi


def mps_overlap(m1: DenseMPS, m2: DenseMPS):
    """<m1|m2> by transfer matrices (any L)."""
    def site_mats(m, i):
        T = m.tensors[i]
        if i == m.ortho_center:
            T = T * m.lams[i][:, None, None]
        if m.ortho_center == m.L and i == m.L - 1:
            T = T * m.lams[m.L][None, None, :]
        return T
    E = np.ones((1, 1), dtype=complex)
    for i in range(m1.L):
        A = site_mats(m1, i).conj()
        B = site_mats(m2, i)
        E = np.einsum("ab,apc,bpd->cd", E, A, B, optimize=True)
    return E[0, 0]


def entropies(lams):
    out = []
    for l in lams:
        p = np.asarray(l) ** 2
        p = p[p > 0]
        out.append(float(-(p * np.log(p)).sum()))
    return np.array(out)



# --------------------------------------------------------------------------------------------
# infinite MPS from two finite chains (slater.py:1356-1565, iMPS.py:65-192)
# --------------------------------------------------------------------------------------------
def basis_rotation(Cm, q_bra, q_ket, S_bra, S_ket, mode="left", form="B"):
    """iMPS.py:65-192 on a dense overlap ``Cm[bra, ket]`` whose charge blocks are q_bra == q_ket
    (``npc.svd`` is block-wise: one SVD per charge sector).  Returns (rotation, unitary_error,
    schmidt_error).  Only the branch used by the mean-field drivers (mode "left") is restated."""
    assert mode == "left"
    S_bra, S_ket = np.asarray(S_bra), np.asarray(S_ket)
    C_Sk = Cm * S_ket[None, :]                                                     # :137
    err2 = np.sum(S_ket ** 2) - np.sum(np.abs(C_Sk) ** 2)                          # :139
    unitary_error = 0.0 if err2 < 0 else float(np.sqrt(err2))                      # :141-155
    if (mode, form) in [("left", "A"), ("right", "B")]:
        M = C_Sk * S_bra[:, None]                                                  # :168
    else:
        M = C_Sk * S_ket[None, :]                                                  # :172
    R = np.zeros_like(Cm)
    for q in np.intersect1d(np.unique(q_bra), np.unique(q_ket)):
        r, c = np.flatnonzero(q_bra == q), np.flatnonzero(q_ket == q)
        U, _, Vh = np.linalg.svd(M[np.ix_(r, c)], full_matrices=False)
        R[np.ix_(r, c)] = U @ Vh                                                   # :173
    if (mode, form) in [("left", "A"), ("right", "B")]:
        Sb_C = R * S_bra[:, None]
    else:
        Sb_C = R * S_ket[None, :]                                                  # :182
    return R, unitary_error, float(np.linalg.norm(Sb_C - C_Sk))                    # :184


@dataclass
class DenseIMPS:
    tensors: list          # T[vL, p, vR], right-canonical ("B" form)
    lams: list             # cell + 1 normalised Schmidt vectors (lams[i] left of site i)
    charges: list          # cell + 1 int arrays (fermion number left of the bond, minus offset)
    errors: tuple          # (left_unitary, left_schmidt, 0.0, 0.0)
    qtotal: list           # charge carried by every tensor


def C_to_iMPS(C_short, C_long, trunc, sites_per_cell, cut, spinful=None, offset="auto") -> DenseIMPS:
    """slater.py:1356-1565 with dense tensors."""
    trunc = Trunc.make(trunc)
    if spinful == "simple":                                                        # :1456-1471
        offset = 2 * round(np.trace(C_short[:cut, :cut]).real) if offset == "auto" else 2 * offset
        C_short, C_long = spinful_correlation_matrix(C_short, False), spinful_correlation_matrix(C_long, False)
        sites_per_cell, cut = 2 * sites_per_cell, 2 * cut
    elif spinful == "PH":
        C_short, C_long = spinful_correlation_matrix(C_short, True), spinful_correlation_matrix(C_long, True)
        sites_per_cell, cut = 2 * sites_per_cell, 2 * cut
    elif spinful is not None:
        raise ValueError("`spinful` must be 'simple', 'PH', or `None`")
    assert len(C_short) + sites_per_cell == len(C_long)
    if offset == "auto":
        offset = round(np.trace(C_short[:cut, :cut]).real)                         # :1491

    def unit(v):
        return v / np.linalg.norm(v)

    short = bond_vectors_from_C(C_short, cut, trunc)                               # :1499-1505
    long_ = bond_vectors_from_C(C_long, cut, trunc)
    lams, tensors, charges, qtot = [unit(short.lam)], [], [long_.n_left - offset], []
    prev = long_
    for i in range(sites_per_cell):                                                # :1509-1537
        if i == sites_per_cell - 1:
            new = short
            lams.append(lams[0])
        else:
            new = bond_vectors_from_C(C_long, cut + i + 1, trunc, "R")
            lams.append(unit(new.lam))
        td = tensor_data(new, prev, "right")
        tensors.append(np.transpose(dense_tensor(td), (2, 0, 1)))
        charges.append(new.n_left - offset)
        qtot.append(td.qtotal)
        prev = new
    td = tensor_data(short, long_, "left")                                         # :1540 (no physical leg)
    R, uerr, serr = basis_rotation(dense_tensor(td), short.n_left, long_.n_left, short.lam, long_.lam)
    tensors[0] = np.tensordot(R, tensors[0], axes=(1, 0))                          # :1554
    charges[0] = short.n_left - offset
    return DenseIMPS(tensors=tensors, lams=lams, charges=charges, errors=(uerr, serr, 0.0, 0.0), qtotal=qtot)


def imps_expectation(imps: DenseIMPS, ops):
    """<prod_i O_i> over one unit cell of a right-canonical iMPS: ``ops`` = list of 2x2 matrices
    (one per site; Jordan-Wigner strings are the caller's business)."""
    lam0 = np.asarray(imps.lams[0])
    E = np.diag(lam0 ** 2).astype(complex)
    for T, O in zip(imps.tensors, ops):
        E = np.einsum("ab,apc,pq,bqd->cd", E, T.conj(), O, T, optimize=True)
    return np.trace(E)


def hopping_chain(L, t=-1.0):
    H = np.zeros((L, L))
    i = np.arange(L - 1)
    H[i, i + 1] = H[i + 1, i] = t
    return H
