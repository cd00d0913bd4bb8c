"""Stage-level parity of the CUDA kernels (through the C ABI) against NumPy / the oracle."""
import ctypes as C

import numpy as np
import pytest

import slater_oracle as so
from temfpy_b200 import _lib
from tests import helpers

pytestmark = pytest.mark.gpu


def _gemm(be, jobs):
    lib = be.lib
    arr = (_lib.GemmJob * len(jobs))(*jobs)
    desc = be.empty(lib.tmf_gemm_desc_bytes(len(jobs)), np.uint8)
    _lib.check(lib, lib.tmf_gemm_grouped(arr, len(jobs), be.ptr(desc), be.stream))
    be.sync()


@pytest.mark.parametrize("tA,tB", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_gemm_grouped_dmma(gpu_backend, tA, tB):
    """FP64 DMMA tiles vs NumPy (tolerance 1e-13 * K: plain FP64 accumulation order differs)."""
    be = gpu_backend
    rng = np.random.default_rng(10 * tA + tB)
    jobs, keep, want = [], [], []
    for (M, N, K) in [(64, 64, 16), (70, 45, 33), (257, 130, 301), (5, 3, 2), (128, 64, 1000)]:
        A = rng.normal(size=(K, M) if tA else (M, K)); B = rng.normal(size=(N, K) if tB else (K, N))
        Cc = rng.normal(size=(M, N))
        dA = be.from_host(np.asfortranarray(A).ravel(order="F")); dB = be.from_host(np.asfortranarray(B).ravel(order="F"))
        dC = be.from_host(np.asfortranarray(Cc).ravel(order="F"))
        keep += [dA, dB, dC]
        jobs.append(_lib.GemmJob(A=be.ptr(dA), B=be.ptr(dB), C=be.ptr(dC), M=M, N=N, K=K, lda=A.shape[0],
                                 ldb=B.shape[0], ldc=M, transA=tA, transB=tB, alpha=1.5, beta=0.5))
        want.append((1.5 * (A.T if tA else A) @ (B.T if tB else B) + 0.5 * Cc, dC, M, N, K))
    _gemm(be, jobs)
    for w, dC, M, N, K in want:
        got = be.to_host(dC, M * N).reshape(N, M).T
        assert np.abs(got - w).max() < 1e-13 * max(K, 16)


def test_corr_build(gpu_backend):
    be = gpu_backend
    L = 300
    H = helpers.random_hamiltonian(L, 1)
    w, v = np.linalg.eigh(H)
    Phi = np.ascontiguousarray(v[:, w < 0])
    d = be.from_host(Phi.ravel()); out = be.empty(L * L, np.float64)
    _lib.check(be.lib, be.lib.tmf_corr_build(be.ptr(d), L, Phi.shape[1], Phi.shape[1], be.ptr(out), L, be.stream))
    be.sync()
    assert np.abs(be.to_host(out, L * L).reshape(L, L) - Phi @ Phi.T).max() < 1e-14


@pytest.mark.parametrize("L,i,chi", [(48, 24, 64), (64, 33, 64), (200, 100, 256), (200, 40, 128)])
def test_minors_blocks_vs_tensor_block(gpu_backend, L, i, chi):
    """K10: same sometimes matrix and occupation masks as the oracle -> every charge block equals
    slater._tensor_block (batched LAPACK det) to 1e-13 (absolute; entries are O(1) or smaller)."""
    be = gpu_backend
    lib = be.lib
    Cm, _ = so.correlation_matrix(helpers.random_hamiltonian(L, L + i) if L < 100 else so.hopping_chain(L))
    tp = {"chi_max": chi}
    oc = L // 2
    if i >= oc:
        ket = so.bond_vectors_from_C(Cm, i, tp, "LR" if i == oc else "R")
        bra = so.bond_vectors_from_C(Cm, i + 1, tp, "R")
        td = so.tensor_data(bra, ket, "right")
    else:
        ket = so.bond_vectors_from_C(Cm, i + 1, tp, "L")
        bra = so.bond_vectors_from_C(Cm, i, tp, "L")
        td = so.tensor_data(bra, ket, "left")
    sb, sk = td.S.shape
    pack = lambda sets: (sets.astype(np.uint64) << np.arange(sets.shape[1], dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint64)
    dS = be.from_host(np.asfortranarray(td.S).ravel(order="F"))
    dbm = be.from_host(pack(td.sets_bra).view(np.int64)); dkm = be.from_host(pack(td.sets_ket).view(np.int64))
    ddet = be.from_host(np.array([float(td.det_always)]))
    qc = 1 if td.mode == "left" else -1
    blocks, outs, want = [], [], []
    for q in np.unique(td.q_ket):
        kr = np.flatnonzero(td.q_ket == q); br = np.flatnonzero(td.q_bra == q + td.qtotal * qc)
        if not br.size:
            continue
        out = be.empty(br.size * kr.size, np.float64)
        blocks.append(_lib.MinorBlock(S=be.ptr(dS), det=be.ptr(ddet), bra_masks=be.ptr(dbm) + 8 * int(br[0]),
                                      ket_masks=be.ptr(dkm) + 8 * int(kr[0]), out=be.ptr(out), s_bra=sb, s_ket=sk,
                                      n_bra=br.size, n_ket=kr.size, minor=int(td.sets_ket[kr[0]].sum())))
        outs.append(out)
        want.append(td.det_always * so.tensor_block(td.S, td.sets_bra[br], td.sets_ket[kr]))
    arr = (_lib.MinorBlock * len(blocks))(*blocks)
    desc = be.empty(lib.tmf_minor_desc_bytes(len(blocks)), np.uint8)
    _lib.check(lib, lib.tmf_minors_blocks(arr, len(blocks), be.ptr(desc), be.stream))
    be.sync()
    for o, w in zip(outs, want):
        assert np.abs(be.to_host(o, w.size).reshape(w.shape) - w).max() < 1e-13


@pytest.mark.parametrize("L,seed", [(40, 1), (150, 5), (300, 8)])
def test_modes_vs_eigh(gpu_backend, L, seed):
    """K3: entangled eigenvalues equal to numpy.linalg.eigh to a few 1e-15 (absolute), mode counts
    identical, eigen-residuals and orthonormality of [entangled | filled] at 1e-13."""
    be = gpu_backend
    lib = be.lib
    Cm, n = so.correlation_matrix(helpers.random_hamiltonian(L, seed))
    tr = so.Trunc(chi_max=64)
    jobs = [(x, s) for x in range(0, L + 1, max(1, L // 23)) for s in (0, 1)]
    jx = (C.c_int * len(jobs))(*[j[0] for j in jobs]); js = (C.c_int * len(jobs))(*[j[1] for j in jobs])
    sizes = [(x if s == 0 else L - x) for x, s in jobs]
    voff = np.concatenate(([0], np.cumsum([m * m + 32 for m in sizes])))[:-1].astype(np.int64)
    V = be.empty(int(sum(m * m + 32 for m in sizes)), np.float64)
    e = be.empty(64 * len(jobs), np.float64); info = be.empty(4 * len(jobs), np.int32)
    wb = lib.tmf_slater_modes_workspace(L, len(jobs), jx, js, 64)
    work = be.empty(wb, np.uint8)
    dC = be.from_host(Cm.ravel())
    _lib.check(lib, lib.tmf_slater_modes_batched(be.ptr(dC), L, L, len(jobs), jx, js, 1e-12, 64,
                                                 voff.ctypes.data_as(_lib.c_i64_p), be.ptr(V), be.ptr(e),
                                                 be.ptr(info), be.ptr(work), wb, be.stream))
    be.sync()
    eh = be.to_host(e).reshape(-1, 64); ih = be.to_host(info).reshape(-1, 4); Vh = be.to_host(V)
    for j, (x, s) in enumerate(jobs):
        mo = so.bond_modes(Cm, x, tr, "LR"[s])
        k, f, m = int(ih[j, 0]), int(ih[j, 1]), sizes[j]
        assert ih[j, 2] == 0
        assert k == mo.e.size and f == mo.n_filled("LR"[s]), (x, s)
        if k:
            assert np.abs(eh[j, :k] - mo.e).max() < 2e-14
        if m == 0:
            continue
        Vm = Vh[voff[j]: voff[j] + m * m].reshape(m, m).T[:, : k + f]
        A = Cm[:x, :x] if s == 0 else Cm[x:, x:]
        eig = np.concatenate([eh[j, :k] if s == 0 else 1 - eh[j, :k], np.ones(f)])
        if k + f == 0:
            continue
        # entangled modes within ~cutoff of the filled cluster are only defined up to eps/gap
        assert np.abs(Vm.T @ Vm - np.eye(k + f)).max() < 1e-8
        if k:
            # sketch path: a mode with singular value s is resolved to ~eps * s_max / s (<= 1e-9 at the
            # cutoff s ~ 1e-6..1e-7); its eigenvalue is second order in that error
            assert np.abs(A @ Vm[:, :k] - Vm[:, :k] * eig[:k]).max() < 2e-9
        if f:
            assert np.abs(Vm[:, k:].T @ A @ Vm[:, k:] - np.eye(f)).max() < 1e-10   # filled space: A = 1
