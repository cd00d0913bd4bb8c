r"""Gutzwiller projection of Abrikosov-fermion MPS onto spin-1/2 MPS (drop-in for ``temfpy.gutzwiller``,
reference gutzwiller.py:22-486).

The reference delegates all arithmetic to TeNPy: ``mps.group_sites(2)`` (gutzwiller.py:227 / :409)
contracts the tensors of the fermion sites :math:`2i, 2i+1`, ``B.iproject`` (:242 / :424) then throws
most of the result away.  Here only the charge-block chains that survive the three masks are ever
multiplied, each as one job of the grouped FP64 tensor-core GEMM (``tmf_gemm_grouped``), batched over
all spin sites in a single launch; the operands are addressed in place inside the block-sparse fermion
tensors (the two physical values of a block are contiguous row ranges).

Input: the :class:`~temfpy_b200.mps.BlockMPS` returned by ``slater.C_to_MPS(..., spinful=...)``
(finite, ``conserve="N"``).  Output: a ``BlockMPS`` with ``site_type="SpinHalfSite"``:

* :func:`abrikosov`     -- ``n = (1,0) -> up``, ``(0,1) -> down``; no charges kept (``conserve=None``,
  physical index 0 = up, 1 = down);
* :func:`abrikosov_ph`  -- ``(0,0) -> down``, ``(1,1) -> up``; ``2 S^z`` conserved (``conserve="Sz"``,
  physical index 0 = down (charge -1), 1 = up (+1), the charge-sorted order of TeNPy's SpinHalfSite;
  virtual charge = fermion number - bond index, gutzwiller.py:437-441).

``return_canonical=True`` brings the result to right-canonical form like ``canonical_form_finite``
(gutzwiller.py:266 / :471): QR and SVD sweeps over the (small, chi_proj ~ chi/2) projected tensors, block-wise in
the charges -- on the host by default, on the device with ``CANONICAL_FORM = "device"`` (``tmf_canon_*``, see there).
Input may conserve the fermion number (Slater) or the parity (Pfaffian, complex tensors), and may be finite or the
unit cell of an infinite MPS (``slater.C_to_iMPS``; ``q_left`` / ``parity`` / ``offset`` as in the reference, canonical
form by the fixed points of the cell transfer matrix).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from warnings import warn

import numpy as np

from . import _lib
from ._lib import check
from .mps import BlockMPS

logger = logging.getLogger(__name__)

CANONICAL_FORM = "host"
"""Where ``return_canonical=True`` runs the QR / SVD sweeps of ``canonical_form_finite``: ``"host"`` (block-wise LAPACK,
the default) or ``"device"`` (``tmf_canon_*``: the sweeps on the projected tensors in HBM, real tensors, charge blocks
up to 160).  The sweep is sequential in the sites and its blocks are small, so a step is bound by the latency of one
CTA's dependent Gram-Schmidt / Jacobi phases: measured on a B200 for BASELINE configs[2] (256 spin sites, blocks <= 67)
the device sweep takes 0.44 s against 0.34 s on the host -- hence the default.  ``TMF_CANONICAL_FORM`` overrides."""


def parity_mask(charges, parity: int = 0) -> np.ndarray:
    """Indices of a leg whose (flat) charge has the given parity (gutzwiller.py:22-48)."""
    return (np.asarray(charges).ravel() % 2 == parity % 2)


def number_mask(charges, n: int) -> np.ndarray:
    """Indices of a leg whose (flat) charge equals ``n`` (gutzwiller.py:51-70)."""
    return (np.asarray(charges).ravel() == n)


@dataclass
class DenseSite:
    """Spin site tensor ``T[vL, p, vR]`` with the flat charges of its legs."""
    T: np.ndarray
    qtotal: int = 0

    def dense(self):
        return self.T


def _check_unit_cell_width(mps: BlockMPS, unit_cell_width, group=2):
    """gutzwiller.py:73-88."""
    if unit_cell_width is None:
        unit_cell_width = mps.unit_cell_width
        if (mps.L // group) % unit_cell_width != 0:
            warn(f"Input MPS {unit_cell_width = } does not divide new MPS size {mps.L // group}\n"
                 "Default to chain geometry")
            unit_cell_width = mps.L // group
    elif (mps.L // group) % unit_cell_width != 0:
        raise ValueError(f"{unit_cell_width = } does not divide new MPS size {mps.L // group}")
    return unit_cell_width


def _unwrap(mps):
    """The BlockMPS behind a TeNPy MPS produced by ``BlockMPS.to_tenpy`` (the drivers return TeNPy objects
    automatically where TeNPy is installed; the projection works on the block storage)."""
    inner = getattr(mps, "_temfpy_b200", None)
    if inner is not None:
        return inner
    if not isinstance(mps, BlockMPS):
        raise TypeError("gutzwiller: expected the result of temfpy_b200.slater.C_to_MPS / H_to_MPS "
                        f"(a BlockMPS, or the TeNPy MPS made from it by to_tenpy()), got {type(mps).__name__}")
    return mps


def _validate(mps):
    assert mps.L % 2 == 0, "Odd-length MPS cannot represent an Abrikosov fermion Hilbert space"
    assert mps.site_type == "FermionSite", f"All sites must be fermionic, found: {mps.site_type}"
    if mps.conserve not in ("N", "parity"):
        raise ValueError(f"FermionSite must conserve either 'N' or 'parity', found {mps.conserve!r}")
    if mps.bc not in ("finite", "infinite"):
        raise NotImplementedError(f"Boundary condition {mps.bc!r} not supported")          # gutzwiller.py:207-208


@dataclass
class _Blk:
    """One charge block ``T[vL in sector qL, p, vR in sector qR]`` of a fermion site tensor: index ranges on the two
    bonds, where its elements live (``("dev", buffer, element offset)`` inside the HBM-resident result of the
    conversion, or ``("host", array)`` to be staged) and the element strides along vL / vR."""
    qL: int
    p: int
    qR: int
    vL0: int
    nL: int
    vR0: int
    nR: int
    src: tuple
    sL: int
    sR: int
    cplx: bool = False


def _contiguous(idx):
    return idx.size > 0 and int(idx[-1]) - int(idx[0]) + 1 == idx.size


def _blocks_of(mps, i, mod):
    """Charge blocks of fermion site ``i`` keyed by ``(qL, p)``.  Results of the Slater conversion are addressed in
    place (left-canonical tensors store [vL, vR] rows, right-canonical ones [vR, vL]; the rows of one physical value
    are a contiguous range of a block, slater.py:1053-1058); any other tensor type is cut out of its dense form."""
    from .engine import SiteTensor
    t = mps.tensors[i]
    out = {}
    if isinstance(t, SiteTensor):
        rp, ra = t.row_p, t.row_alpha
        left = t.mode == "left"
        cplx = bool(t.blocks) and np.iscomplexobj(t.blocks[0][5])
        for bi, (qk, r0, nr, c0, nc, arr) in enumerate(t.blocks):
            # the rows of a block are [p = 0 | p = 1] (stable sort by pipe charge), each a contiguous range of alpha
            n0 = int(np.searchsorted(rp[r0: r0 + nr], 1))
            for p, s0, n in ((0, 0, n0), (1, n0, nr - n0)):
                if n == 0:
                    continue
                a0 = int(ra[r0 + s0])
                assert int(ra[r0 + s0 + n - 1]) == a0 + n - 1
                src = ("dev", t.dev[0], t.dev[1][bi] + s0 * nc) if t.dev is not None else ("host", arr[s0: s0 + n])
                if left:
                    out[(qk - p, p)] = _Blk(qk - p, p, qk, a0, n, c0, nc, src, nc, 1, cplx)
                else:
                    out[(qk, p)] = _Blk(qk, p, qk + p, c0, nc, a0, n, src, 1, nc, cplx)
        return out
    T = t.dense()
    cL, cR = np.asarray(mps.charges[i]).ravel(), np.asarray(mps.charges[i + 1]).ravel()
    qt = int(getattr(t, "qtotal", 0) or 0)          # q(vL) + p - q(vR) (unit cells: the last tensor carries the filling)
    for qL in np.unique(cL):
        iL = np.flatnonzero(cL == qL)
        for p in (0, 1):
            qR = int(qL) + p - qt
            if mod:
                qR %= mod
            iR = np.flatnonzero(cR == qR)
            if iR.size == 0:
                continue
            if not (_contiguous(iL) and _contiguous(iR)):
                raise NotImplementedError("gutzwiller: charge sectors of the input MPS must be contiguous index ranges")
            sub = np.ascontiguousarray(T[iL[0]: iL[-1] + 1, p, iR[0]: iR[-1] + 1])
            if not sub.any():
                continue
            out[(int(qL), p)] = _Blk(int(qL), p, qR, int(iL[0]), iL.size, int(iR[0]), iR.size, ("host", sub),
                                     iR.size, 1, np.iscomplexobj(sub))
    return out


def _project(mps: BlockMPS, keep, spin_rules, be, mod=0):
    """All pairs (2j, 2j+1) contracted and projected in one launch of ``tmf_gutzwiller_project``.

    keep(j, charges) -> bool mask of the virtual indices of bond 2j that survive; spin_rules = ((s, p_a, p_b), ...)
    the pair occupations that become spin index s.  Only the charge chains (qL, p_a) -> q_m -> (p_b, qR) with
    surviving ends are multiplied; the operands are read in place in HBM when the fermion MPS is resident there
    (otherwise the needed blocks are staged once), the results land at their place in the dense spin-site tensors
    ``T[vL, s, vR]``.  Returns (tensors on the host, kept indices per spin bond, number of chains, device copy)."""
    lib = be.lib
    Ls = mps.L // 2
    finite = mps.bc == "finite"
    oc = mps.ortho_center if finite else None
    nb = Ls + 1 if finite else Ls
    bond = lambda j: 2 * j if finite else (2 * j) % mps.L
    keepers = [np.flatnonzero(keep(j, np.asarray(mps.charges[bond(j)]).ravel())) for j in range(nb)]
    pos = []
    for j in range(nb):
        m = -np.ones(len(np.asarray(mps.charges[bond(j)]).ravel()), dtype=np.int64)
        m[keepers[j]] = np.arange(len(keepers[j]))
        pos.append(m)
    # chains
    chains, staged, stage_off = [], [], 0
    out_off, offs = 0, []
    for j in range(Ls):
        jr = j + 1 if finite else (j + 1) % Ls
        nLk, nRk = len(keepers[j]), len(keepers[jr])
        offs.append(out_off)
        out_off += nLk * 2 * nRk
        if nLk == 0 or nRk == 0:
            continue
        ba, bb = _blocks_of(mps, 2 * j, mod), _blocks_of(mps, 2 * j + 1, mod)
        for (s, pa, pb) in spin_rules:
            for (qL, p), A in ba.items():
                if p != pa or pos[j][A.vL0] < 0:
                    continue
                B = bb.get((A.qR, pb))
                if B is None or pos[jr][B.vR0] < 0:
                    continue
                assert (A.vR0, A.nR) == (B.vL0, B.nL), "bond sectors of neighbouring tensors do not match"
                chains.append((j, s, A, B, int(pos[j][A.vL0]), int(pos[jr][B.vR0]), nRk))
    cplx = any(X.cplx for c in chains for X in (c[2], c[3]))
    es = 2 if cplx else 1
    dt = np.complex128 if cplx else np.float64
    for (_, _, A, B, *_r) in chains:
        for X in (A, B):
            if X.src[0] == "host" and len(X.src) == 2:
                arr = np.ascontiguousarray(X.src[1], dtype=dt).ravel()
                staged.append(arr)
                X.src = ("host", X.src[1], stage_off)
                stage_off += arr.size
    stage_d = be.from_host(np.concatenate(staged).view(np.float64)) if staged else None
    lam_d = None
    if oc is not None:
        lam_d = be.from_host(np.ascontiguousarray(mps.lams[oc], dtype=np.float64))
    outd = be.empty(max(es * out_off, 1), np.float64)

    def addr(X):
        if X.src[0] == "dev":
            return be.ptr(X.src[1]) + 8 * es * int(X.src[2])
        return be.ptr(stage_d) + 8 * es * int(X.src[2])
    jobs = (_lib.GutzJob * max(len(chains), 1))()
    for u, (j, s, A, B, pL, pR, nRk) in enumerate(chains):
        g = jobs[u]
        g.A, g.sa_i, g.sa_k = addr(A), A.sL, A.sR
        g.B, g.sb_k, g.sb_n = addr(B), B.sL, B.sR
        g.m, g.k, g.n = A.nL, A.nR, B.nR
        g.out = be.ptr(outd) + 8 * es * (offs[j] + (pL * 2 + s) * nRk + pR)
        g.so_i = 2 * nRk
        if oc is not None:
            if oc == 2 * j + 1:                 # Schmidt values of the centre bond sit inside this pair
                g.k_scale = be.ptr(lam_d) + 8 * A.vR0
            elif oc == 2 * j:                   # ... on its left bond: A..A lam B..B
                g.row_scale = be.ptr(lam_d) + 8 * A.vL0
            elif oc == mps.L and j == Ls - 1:
                g.col_scale = be.ptr(lam_d) + 8 * B.vR0
    desc = be.empty(int(lib.tmf_gutz_desc_bytes(jobs, len(chains))), np.uint8)
    check(lib, lib.tmf_gutzwiller_project(jobs, len(chains), int(cplx), be.ptr(outd), 8 * es * out_off,
                                          be.ptr(desc), be.stream))
    be.sync()
    outs = be.to_host(outd, es * out_off)
    if cplx:
        outs = outs.view(np.complex128)
    tensors = []
    for j in range(Ls):
        jr = j + 1 if finite else (j + 1) % Ls
        nLk, nRk = len(keepers[j]), len(keepers[jr])
        tensors.append(outs[offs[j]: offs[j] + nLk * 2 * nRk].reshape(nLk, 2, nRk))
    resident = sum(1 for c in chains for X in (c[2], c[3]) if X.src[0] == "dev")
    return tensors, keepers, dict(chains=len(chains), resident_operands=resident, staged_elems=stage_off,
                                  device=(outd, offs), cplx=cplx)


def _svd(M):
    """LAPACK gesdd with the gesvd fallback (gesdd occasionally fails to converge on graded matrices)."""
    try:
        return np.linalg.svd(M, full_matrices=False)
    except np.linalg.LinAlgError:
        from scipy.linalg import svd
        return svd(M, full_matrices=False, lapack_driver="gesvd")


def _canonical_form_finite(tensors, qs, qp, cutoff):
    """Right-canonical form of a finite MPS given by bare tensors ``T[vL, p, vR]`` (the job of
    ``canonical_form_finite`` at gutzwiller.py:266 / :471), block-wise in the charges: ``qs[j]`` flat charges
    of bond j, ``qp`` of the physical index, with ``q(vL) + qp = q(vR)``.  Returns (tensors, lams, charges)."""
    L = len(tensors)
    T = [t.copy() for t in tensors]
    qs = [np.asarray(q).copy() for q in qs]
    qp = np.asarray(qp)
    # left-to-right QR sweep -> left-canonical
    for j in range(L - 1):
        a, d, b = T[j].shape
        rowq = (qs[j][:, None] + qp[None, :]).ravel()
        M = T[j].reshape(a * d, b)
        newQ, newR, newq = [], [], []
        for q in np.unique(qs[j + 1]):
            r, c = np.flatnonzero(rowq == q), np.flatnonzero(qs[j + 1] == q)
            if r.size == 0:
                continue
            Q, R = np.linalg.qr(M[np.ix_(r, c)])
            newQ.append((r, Q))
            newR.append((c, R))
            newq += [q] * Q.shape[1]
        k = len(newq)
        Qf, Rf = np.zeros((a * d, k), dtype=M.dtype), np.zeros((k, b), dtype=M.dtype)
        o = 0
        for (r, Q), (c, R) in zip(newQ, newR):
            w = Q.shape[1]
            Qf[r, o: o + w] = Q
            Rf[o: o + w, c] = R
            o += w
        T[j] = Qf.reshape(a, d, k)
        T[j + 1] = np.tensordot(Rf / np.linalg.norm(Rf), T[j + 1], axes=(1, 0))   # the norm is fixed at the end
        qs[j + 1] = np.array(newq, dtype=np.int64)
    # right-to-left SVD sweep -> right-canonical + Schmidt values
    lams = [None] * (L + 1)
    nrm = np.linalg.norm(T[-1])
    T[-1] = T[-1] / nrm
    lams[L] = np.ones(T[-1].shape[2])
    for j in range(L - 1, -1, -1):
        a, d, b = T[j].shape
        colq = (qs[j + 1][None, :] - qp[:, None]).ravel()
        M = T[j].reshape(a, d * b)
        parts, newq = [], []
        for q in np.unique(qs[j]):
            r, c = np.flatnonzero(qs[j] == q), np.flatnonzero(colq == q)
            if c.size == 0:
                continue
            U, S, Vh = _svd(M[np.ix_(r, c)])
            keep = S > cutoff
            parts.append((r, c, U[:, keep], S[keep], Vh[keep]))
            newq += [q] * int(keep.sum())
        k = len(newq)
        Uf, Sf, Vf = np.zeros((a, k), dtype=M.dtype), np.zeros(k), np.zeros((k, d * b), dtype=M.dtype)
        o = 0
        for r, c, U, S, Vh in parts:
            w = len(S)
            Uf[r, o: o + w] = U
            Sf[o: o + w] = S
            Vf[o: o + w, c] = Vh
            o += w
        Sf = Sf / np.linalg.norm(Sf)
        T[j] = Vf.reshape(k, d, b)
        lams[j] = Sf
        if j > 0:
            T[j - 1] = np.tensordot(T[j - 1], Uf * Sf[None, :], axes=(2, 0))
            qs[j] = np.array(newq, dtype=np.int64)
        else:
            qs[0] = np.array(newq, dtype=np.int64)
    return T, lams, qs


def _canonical_form_device(be, dev, qvirt, qp, cutoff):
    """``canonical_form_finite`` (gutzwiller.py:266 / :471) on the device: ``tmf_canon_*`` plans the QR and SVD sweeps
    from the sector tables and enqueues them on the projected tensors where ``tmf_gutzwiller_project`` left them
    (three launches per site and sweep, no host synchronisation); only the final right-canonical tensors and the
    singular values come back.  Singular values <= cutoff are removed here, at the end (equivalent to the
    reference's truncation on the fly: a discarded direction only multiplies zeros afterwards).
    Returns (tensors, lams, charges) or None if a charge block exceeds the kernels' limits."""
    import ctypes as C
    lib = be.lib
    outd, offs = dev
    L = len(offs)
    dims0 = np.array([len(q) for q in qvirt], dtype=np.int32)
    ch0 = np.ascontiguousarray(np.concatenate([np.asarray(q) for q in qvirt]).astype(np.int32))
    qpi = np.ascontiguousarray(np.asarray(qp, dtype=np.int32))
    ip = lambda a: a.ctypes.data_as(_lib.c_int_p)
    h = lib.tmf_canon_create(L, ip(dims0), ip(ch0), ip(qpi))
    if not h:
        logger.info("device canonical form not applicable (%s): host sweep", lib.tmf_last_error().decode())
        return None
    try:
        q = (C.c_int64 * 8)()
        check(lib, lib.tmf_canon_sizes(h, q))
        wb, nt, ns, nd = int(q[0]), int(q[1]), int(q[2]), int(q[3])
        dims2 = np.zeros(L + 1, dtype=np.int32)
        ch2 = np.zeros(max(nd, 1), dtype=np.int32)
        check(lib, lib.tmf_canon_dims(h, ip(dims2), ip(ch2)))
        work = be.empty(wb, np.uint8)
        T2 = be.empty(max(nt, 1), np.float64)
        S = be.empty(max(ns, 1), np.float64)
        inv = be.empty(L + 2, np.float64)
        t0 = (C.c_int64 * L)(*[int(o) for o in offs])
        check(lib, lib.tmf_canon_run(h, be.ptr(outd), t0, be.ptr(work), wb, be.ptr(T2), be.ptr(S), be.ptr(inv),
                                     be.stream))
        be.sync()
        T2h, Sh = be.to_host(T2, max(nt, 1)), be.to_host(S, max(ns, 1))
    finally:
        lib.tmf_canon_destroy(h)
    boff = np.concatenate(([0], np.cumsum(dims2))).astype(np.int64)
    keep, lams, charges = [], [], []
    for j in range(L + 1):
        d = int(dims2[j])
        if j < L:
            s = np.array(Sh[boff[j]: boff[j] + d])
            k = s > cutoff
            sk = s[k]
            lams.append(sk / np.linalg.norm(sk) if sk.size else sk)
        else:
            k = np.ones(d, dtype=bool)
            lams.append(np.ones(d))
        keep.append(k)
        charges.append(ch2[boff[j]: boff[j] + d][k].astype(np.int64))
    tensors, o = [], 0
    for j in range(L):
        a, b = int(dims2[j]), int(dims2[j + 1])
        T = T2h[o: o + a * 2 * b].reshape(a, 2, b)
        o += a * 2 * b
        if not keep[j].all():
            T = T[keep[j]]
        if not keep[j + 1].all():
            T = T[:, :, keep[j + 1]]
        tensors.append(T)
    return tensors, lams, charges


def _infer_qtot(tensors, qs, qp):
    """q(vL) + qp[s] - q(vR) of every tensor of a unit cell (constant over its non-zero entries)."""
    out = []
    L = len(tensors)
    for j, t in enumerate(tensors):
        a, p, b = np.nonzero(np.abs(t) > 0)
        if a.size == 0:
            out.append(0)
            continue
        d = np.unique(np.asarray(qs[j])[a] + np.asarray(qp)[p] - np.asarray(qs[(j + 1) % L])[b])
        assert d.size == 1, f"spin tensor {j} mixes total charges {d}"
        out.append(int(d[0]))
    return out


def _sectors(q):
    q = np.asarray(q)
    return [(int(v), np.flatnonzero(q == v)) for v in np.unique(q)]


def _cell_map_right(T, r):
    """r (bond 0 of the next cell) -> sum over the cell of T r T^+ (bond 0)."""
    for t in reversed(T):
        r = np.einsum("apb,bc,dpc->ad", t, r, t.conj(), optimize=True)
    return r


def _cell_map_left(T, l):
    for t in T:
        l = np.einsum("apb,ac,cpd->bd", t.conj(), l, t, optimize=True)
    return l


def _dominant(apply, n, dtype):
    """Dominant eigenpair of a completely positive map on n x n matrices, started from the identity (so that the
    iterates keep the block structure of the charge sectors exactly): Arnoldi, power iteration as a fallback."""
    v0 = np.eye(n, dtype=dtype).ravel()
    try:
        from scipy.sparse.linalg import LinearOperator, eigs
        if n * n <= 4:
            raise ValueError
        op = LinearOperator((n * n, n * n), matvec=lambda v: apply(v.reshape(n, n)).ravel(), dtype=np.complex128)
        w, v = eigs(op, k=1, which="LM", v0=v0.astype(np.complex128), tol=1e-14, maxiter=5000)
        eta, m = w[0], v[:, 0].reshape(n, n)
    except Exception:
        m = v0.reshape(n, n)
        eta = 1.0
        for _ in range(20000):
            m2 = apply(m)
            eta2 = np.linalg.norm(m2)
            m2 = m2 / eta2
            if np.linalg.norm(m2 - m) < 1e-15:
                m, eta = m2, eta2
                break
            m, eta = m2, eta2
    m = m / (np.trace(m) / abs(np.trace(m)))          # Hermitian positive up to a phase
    m = (m + m.conj().T) / 2
    if np.trace(m).real < 0:
        m = -m
    return abs(eta), (m if np.iscomplexobj(np.zeros(1, dtype=dtype)) else m.real)


def _canonical_form_infinite(tensors, qs, qp, qtot, cutoff=1e-12):
    """Right-canonical form of an infinite MPS given by the bare tensors ``T_j[vL, p, vR]`` of a unit cell (the job of
    ``canonical_form_infinite1`` at gutzwiller.py:268 / :473), block-wise in the charges: ``qs[j]`` flat charges of
    bond j (bond L is bond 0 of the next cell), ``q(vL) + qp[p] - qtot[j] = q(vR)``.

    Dominant left / right eigenvectors l, r of the cell transfer matrix; r = X X^+, l = Y^+ Y per sector, SVD of Y X
    gives the basis of bond 0 in which r = 1 and l = lambda^2; a right-to-left SVD sweep makes the inner tensors
    right-canonical, a left-to-right pass of the left environment diagonalises it on the inner bonds.
    Returns (tensors, lams, charges)."""
    L = len(tensors)
    T = [np.asarray(t) for t in tensors]
    dtype = np.result_type(*[t.dtype for t in T])
    qs = [np.asarray(q) for q in qs[:L]]
    qp = np.asarray(qp)
    n0 = T[0].shape[0]
    eta, r = _dominant(lambda m: _cell_map_right(T, m), n0, dtype)
    T = [t / eta ** (0.5 / L) for t in T]
    _, l = _dominant(lambda m: _cell_map_left(T, m), n0, dtype)
    # bond 0: per charge sector  r = X X^+,  l = Y^+ Y,  Y X = U lam V^+  ->  G = X V,  G^-1 = V^+ X^-1
    cols, icols, lam0, q0 = [], [], [], []
    for q, idx in _sectors(qs[0]):
        wr, Ur = np.linalg.eigh(r[np.ix_(idx, idx)])
        wl, Ul = np.linalg.eigh(l[np.ix_(idx, idx)])
        kr, kl = wr > 1e-14 * max(wr.max(), 1e-300), wl > 1e-14 * max(wl.max(), 1e-300)
        if not kr.any() or not kl.any():
            continue
        X = Ur[:, kr] * np.sqrt(wr[kr])[None, :]
        Xi = (Ur[:, kr] / np.sqrt(wr[kr])[None, :]).conj().T
        Y = np.sqrt(wl[kl])[:, None] * Ul[:, kl].conj().T
        U, sv, Vh = _svd(Y @ X)
        cols.append((idx, X @ Vh.conj().T))
        icols.append((idx, Vh @ Xi))
        lam0.append(sv)
        q0 += [q] * len(sv)
    lam0 = np.concatenate(lam0)
    keep0 = lam0 > cutoff * lam0.max()
    k0 = int(keep0.sum())
    G, Gi = np.zeros((n0, len(lam0)), dtype=dtype), np.zeros((len(lam0), n0), dtype=dtype)
    o = 0
    for (idx, g), (_, gi) in zip(cols, icols):
        w = g.shape[1]
        G[idx, o: o + w] = g
        Gi[o: o + w, idx] = gi
        o += w
    G, Gi, lam0, q0 = G[:, keep0], Gi[keep0], lam0[keep0], np.array(q0, dtype=np.int64)[keep0]
    lam0 = lam0 / np.linalg.norm(lam0)
    # right-to-left sweep: B_j = Vh of (T_j G_{j+1}), G_j = U S
    newq = [None] * (L + 1)
    newq[0] = newq[L] = q0
    B = [None] * L
    Gr = G
    for j in range(L - 1, 0, -1):
        A = np.tensordot(T[j], Gr, axes=(2, 0))                       # (a_j, 2, k_{j+1})
        a, d, b = A.shape
        colq = (newq[j + 1][None, :] - qp[:, None] + qtot[j]).ravel()
        M = A.reshape(a, d * b)
        parts, nq = [], []
        for q, rows in _sectors(qs[j]):
            c = np.flatnonzero(colq == q)
            if c.size == 0:
                continue
            U, S, Vh = _svd(M[np.ix_(rows, c)])
            kp = S > cutoff * max(S.max(), 1e-300)
            parts.append((rows, c, U[:, kp] * S[kp][None, :], Vh[kp]))
            nq += [q] * int(kp.sum())
        k = len(nq)
        Gn, Bj = np.zeros((a, k), dtype=dtype), np.zeros((k, d * b), dtype=dtype)
        o = 0
        for rows, c, us, vh in parts:
            w = vh.shape[0]
            Gn[rows, o: o + w] = us
            Bj[o: o + w, c] = vh
            o += w
        B[j] = Bj.reshape(k, d, b)
        newq[j] = np.array(nq, dtype=np.int64)
        Gr = Gn
    B[0] = np.tensordot(Gi, np.tensordot(T[0], Gr, axes=(2, 0)), axes=(1, 0)) if L > 1 else \
        np.tensordot(Gi, np.tensordot(T[0], G, axes=(2, 0)), axes=(1, 0))
    # left-to-right: diagonalise the left environment on the inner bonds
    lams = [None] * (L + 1)
    lams[0] = lams[L] = lam0
    lcur = np.diag(lam0 ** 2).astype(dtype)
    for j in range(L - 1):
        ln = np.einsum("apb,ac,cpd->bd", B[j].conj(), lcur, B[j], optimize=True)
        k = ln.shape[0]
        W, w_all = np.zeros((k, k), dtype=dtype), np.zeros(k)
        for q, idx in _sectors(newq[j + 1]):
            w, U = np.linalg.eigh(ln[np.ix_(idx, idx)])
            W[np.ix_(idx, idx)] = U[:, ::-1]
            w_all[idx] = w[::-1]
        w_all = np.clip(w_all, 0.0, None)
        B[j] = np.tensordot(B[j], W, axes=(2, 0))
        B[j + 1] = np.tensordot(W.conj().T, B[j + 1], axes=(1, 0))
        lam = np.sqrt(w_all)
        lams[j + 1] = lam / np.linalg.norm(lam)
        lcur = np.diag(w_all / w_all.sum()).astype(dtype)
    return B, lams, newq


def _finish(mps, tensors, keepers, qvirt, qp, conserve, return_canonical, cutoff, unit_cell_width, info, be=None):
    import os
    Ls = len(tensors)
    meta = dict(gemm_jobs=info["chains"], kept=[len(k) for k in keepers], resident_operands=info["resident_operands"],
                staged_elems=info["staged_elems"])
    if mps.bc == "infinite":
        qtot = _infer_qtot(tensors, qvirt, qp)
        if return_canonical:
            T, lams, qs = _canonical_form_infinite(tensors, qvirt, qp, qtot, max(cutoff, 1e-14))      # :268 / :473
            form = ["B"] * Ls
            meta["canonical_form"] = "host"
            logger.info("Transformed MPS to right canonical form")
        else:
            warn("The MPS is not in canonical form after Gutzwiller projection.\n"
                 "Consider setting 'return_canonical=True'")
            T, qs, form = tensors, list(qvirt) + [qvirt[0]], [None] * Ls
            lams = [np.ones(len(q)) / np.sqrt(max(len(q), 1)) for q in qs]
        return BlockMPS(L=Ls, tensors=[DenseSite(t, int(qt)) for t, qt in zip(T, qtot)], lams=lams, charges=qs,
                        form=form, unit_cell_width=unit_cell_width, ortho_center=None, bc="infinite",
                        site_type="SpinHalfSite", conserve=conserve, meta=meta)
    if return_canonical:
        got = None
        mode = os.environ.get("TMF_CANONICAL_FORM", CANONICAL_FORM)
        if mode == "device" and be is not None and not info["cplx"]:
            got = _canonical_form_device(be, info["device"], qvirt, qp, cutoff)
        meta["canonical_form"] = "device" if got is not None else "host"
        T, lams, qs = got if got is not None else _canonical_form_finite(tensors, qvirt, qp, cutoff)
        form = ["B"] * Ls
        oc = 0
        logger.info("Transformed MPS to right canonical form")
    else:
        warn("The MPS is not in canonical form after Gutzwiller projection.\n"
             "Consider setting 'return_canonical=True'")
        T, qs, form, oc = tensors, qvirt, [None] * Ls, None
        lams = [np.ones(len(q)) / np.sqrt(max(len(q), 1)) for q in qs]                       # gutzwiller.py:258
        meta["device"] = info["device"]             # the projected tensors also stay in HBM: (buffer, offsets)
    return BlockMPS(L=Ls, tensors=[DenseSite(t) for t in T], lams=lams, charges=qs, form=form,
                    unit_cell_width=unit_cell_width, ortho_center=oc, bc="finite", site_type="SpinHalfSite",
                    conserve=conserve, meta=meta)


def _deliver(given, inner, out, inplace):
    """``inplace=True`` (gutzwiller.py:210 / :280): the projected state replaces the contents of the given
    :class:`BlockMPS` and nothing is returned, like the reference does with the TeNPy object.  A TeNPy MPS made by
    ``to_tenpy()`` cannot be rewritten from here (its tensors are TeNPy's): ask for the result instead."""
    if not inplace:
        return out
    if given is not inner:
        raise NotImplementedError("`inplace=True` needs the BlockMPS itself (a TeNPy MPS is not rewritten in place)")
    inner.__dict__.update(out.__dict__)
    return None


def abrikosov(mps: BlockMPS, *, inplace: bool = False, return_canonical: bool = True, cutoff: float = 1e-12,
              q_left: None | int = None, unit_cell_width: int | None = None, _backend=None):
    r"""Projection from Abrikosov fermions to a spin-1/2 Hilbert space (gutzwiller.py:95-281):
    single occupation of :math:`f_{i\uparrow}` -> up, of :math:`f_{i\downarrow}` -> down.  The input may conserve
    the fermion number (``slater.C_to_MPS``) or only the parity (``pfaffian.C_to_MPS``)."""
    from . import slater as _sl
    given = mps
    mps = _unwrap(mps)
    _validate(mps)
    target = mps.L // 2
    if mps.bc == "finite":
        total = int(np.asarray(mps.charges[mps.L]).ravel()[0])
        if q_left not in (None, 0):
            warn(f"`q_left` must be 0 for finite MPS, got {q_left = }, setting it to 0.")
        q_left = 0
    else:                                                                              # gutzwiller.py:197-205
        total = int(sum(int(getattr(t, "qtotal", 0) or 0) for t in mps.tensors))
        if q_left is None:
            raise ValueError("Must specify `q_left` for infinite MPS.")
        sectors0 = np.unique(np.asarray(mps.charges[0]).ravel())
        if q_left not in sectors0:
            raise ValueError(f"`q_left` must be a charge sector of the leftmost virtual leg, got {q_left = }, "
                             f"valid sectors are {sectors0}")
    err = f"Total charge must match number of spin sites. Got {total}, expected {target}"
    if mps.conserve == "N":
        assert total == target, err                                                    # gutzwiller.py:177-186
        mod = 0
        keep = lambda j, q: number_mask(q, q_left + j)   # one fermion per pair: bond 2j carries q_left + j (:236-238)
    else:
        assert total % 2 == target % 2, err + " (mod 2)"
        mod = 2
        keep = lambda j, q: parity_mask(q, q_left + j)
    ucw = _check_unit_cell_width(mps, unit_cell_width)
    be = _backend or _sl._be()
    tensors, keepers, info = _project(mps, keep, ((0, 1, 0), (1, 0, 1)), be, mod)
    qvirt = [np.zeros(len(k), dtype=np.int64) for k in keepers]      # all charges dropped (:244)
    logger.info("Completed projection to spin-1/2 space. No conserved charges left.")
    return _deliver(given, mps, _finish(mps, tensors, keepers, qvirt, np.zeros(2, dtype=np.int64), None,
                                        return_canonical, cutoff, ucw, info, be), inplace)


def abrikosov_ph(mps: BlockMPS, *, inplace: bool = False, return_canonical: bool = True, cutoff: float = 1e-12,
                 offset: int = 0, parity: int = 0, unit_cell_width: int | None = None, _backend=None):
    r"""Projection from particle-hole rotated Abrikosov fermions (gutzwiller.py:284-486):
    zero occupation -> down, double occupation -> up.  For number-conserving input :math:`2S^z` = number - bond
    index is conserved; parity-conserving input (``pfaffian.C_to_MPS``) leaves no charge (:364-367, :444)."""
    from . import slater as _sl
    given = mps
    mps = _unwrap(mps)
    _validate(mps)
    finite = mps.bc == "finite"
    total = int(np.asarray(mps.charges[mps.L]).ravel()[0]) if finite else \
        int(sum(int(getattr(t, "qtotal", 0) or 0) for t in mps.tensors))
    assert total % 2 == 0, f"Total fermion parity of MPS must be even, got {total}"
    if finite:
        if parity != 0:
            warn(f"Must use even parity sector in finite MPS, ignoring {parity = }")
        if offset != 0 and mps.conserve == "N":
            warn(f"Cannot offset charge of finite MPS, ignoring {offset = }")
        offset = parity = 0
    ucw = _check_unit_cell_width(mps, unit_cell_width)
    be = _backend or _sl._be()
    keep = lambda j, q: parity_mask(q, parity)                                               # :416-417
    # (0,0) -> down (index 0, 2Sz = -1), (1,1) -> up (index 1, 2Sz = +1)
    tensors, keepers, info = _project(mps, keep, ((0, 0, 0), (1, 1, 1)), be, 0 if mps.conserve == "N" else 2)
    if mps.conserve == "N":
        nbond = len(keepers)
        bond = (lambda j: 2 * j) if finite else (lambda j: (2 * j) % mps.L)
        qvirt = [np.asarray(mps.charges[bond(j)])[keepers[j]].astype(np.int64) - (offset + j)
                 for j in range(nbond)]                                                      # :437-441
        qp, conserve = np.array([-1, 1], dtype=np.int64), "Sz"
    else:
        qvirt = [np.zeros(len(k), dtype=np.int64) for k in keepers]
        qp, conserve = np.zeros(2, dtype=np.int64), None
    logger.info("Completed projection to spin-1/2 space. Conserved charge is now %s", conserve)
    return _deliver(given, mps, _finish(mps, tensors, keepers, qvirt, qp, conserve, return_canonical, cutoff, ucw, info,
                                        be), inplace)
