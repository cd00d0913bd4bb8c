"""slater.C_to_iMPS (reference slater.py:1356-1565, iMPS.py:65-192): oracle known answers, the device
driver on the CPU simulator and on the GPU."""
import numpy as np
import pytest

import slater_oracle as so
from temfpy_b200 import iMPS, slater


def dimer_chain(L, t1=-1.0, t2=-1.5, tnn=0.0):
    """examples/iMPS_slater.py:6-23 (BASELINE cfg5, iMPS half): dimerised hopping chain.  ``tnn`` adds a
    next-nearest-neighbour hopping that breaks the particle-hole symmetry (and with it the exact degeneracies
    of the Schmidt spectrum, see _compare)."""
    H = np.zeros((L, L))
    for i in range(L - 1):
        H[i, i + 1] = H[i + 1, i] = t1 if i % 2 == 0 else t2
    for i in range(L - 2):
        H[i, i + 2] = H[i + 2, i] = tnn
    return H


def cell_transfer_eig(A, B):
    """dominant eigenvalue of the mixed transfer matrix of two unit cells (lists of T[vL, p, vR])."""
    E = None
    for Ta, Tb in zip(A, B):
        M = np.einsum("apc,bpd->abcd", Ta.conj(), Tb).reshape(Ta.shape[0] * Tb.shape[0], Ta.shape[2] * Tb.shape[2])
        E = M if E is None else E @ M
    w = np.linalg.eigvals(E)
    return w[np.argmax(np.abs(w))]


def _case(Ls, cell, cut, tp, spinful=None, tnn=0.2):
    Cs, _ = so.correlation_matrix(dimer_chain(Ls, tnn=tnn))
    Cl, _ = so.correlation_matrix(dimer_chain(Ls + cell, tnn=tnn))
    return Cs, Cl, so.C_to_iMPS(Cs, Cl, tp, cell, cut, spinful=spinful)


def test_oracle_imps_known_answers():
    """the unit cell reproduces the bulk correlators of the long chain and is (nearly) unitary-gauged."""
    Cs, Cl, im = _case(64, 2, 32, {"chi_max": 40}, tnn=0.0)
    assert im.errors[0] < 1e-3 and im.errors[1] < 1e-3
    n, I = np.diag([0, 1.0]), np.eye(2)
    cd = np.array([[0, 0], [1, 0.0]])
    assert abs(so.imps_expectation(im, [n, I]) - Cl[32, 32]) < 1e-6
    assert abs(so.imps_expectation(im, [I, n]) - Cl[33, 33]) < 1e-6
    assert abs(so.imps_expectation(im, [cd, cd.T]) - Cl[32, 33]) < 1e-5
    assert abs(abs(cell_transfer_eig(im.tensors, im.tensors)) - 1) < 1e-5


def clean_chi(Cs, Cl, cell, cut, target, spinful=None):
    """a chi_max near ``target`` whose cut falls into a clear gap (> 5 % in lambda) of the Schmidt spectrum of
    every bond involved, so that the kept sets do not depend on rounding noise (see _compare)."""
    big = so.C_to_iMPS(Cs, Cl, {"chi_max": target + 24}, cell, cut, spinful=spinful)
    spectra = [np.sort(l)[::-1] for l in big.lams]
    for d in range(0, 20):
        for chi in (target + d, target - d):
            if all(len(l) > chi and l[chi - 1] / l[chi] > 1.05 for l in spectra):
                return chi
    raise RuntimeError(f"no clean cut found near {target}: {[len(l) for l in spectra]}")


def _compare(im_ref, mps, err, err2_tol=1e-13):
    """NB the test chains break particle-hole symmetry (tnn) to keep the chi_max cut away from (near-)degenerate
    Schmidt multiplets: there the kept set is
    decided by the rounding noise of the weakest mode eigenvalues (e ~ 1e-12 known to ~1e-16 absolute),
    in the reference as much as here, and the two chains of a pair may even truncate differently."""
    got = [mps.get_B_dense(i) for i in range(mps.L)]
    assert [len(l) for l in mps.lams] == [len(l) for l in im_ref.lams]
    for a, b in zip(im_ref.lams, mps.lams):      # tolerance model of helpers.compare_mps
        assert np.all(np.abs(a - b) <= 1e-12 * a + np.minimum(1e-13 / (2 * a), 1e-8)), np.abs(a - b).max()
    for a, b in zip(im_ref.charges, mps.charges):
        assert np.array_equal(a, b)
    assert mps.meta["qtotal"] == im_ref.qtotal
    # unitary_error^2 is a difference of O(1) numbers (iMPS.py:139): compare the squares at rounding level
    assert abs(err.left_unitary ** 2 - im_ref.errors[0] ** 2) < err2_tol and abs(err.left_schmidt - im_ref.errors[1]) < 1e-9
    e_mix = cell_transfer_eig(im_ref.tensors, got)
    e_ref, e_got = cell_transfer_eig(im_ref.tensors, im_ref.tensors), cell_transfer_eig(got, got)
    fid = abs(e_mix) ** 2 / abs(e_ref * e_got)
    assert fid >= 1 - 1e-10, fid
    return fid


@pytest.mark.parametrize("Ls,cell,cut,target,spinful", [(32, 2, 16, 40, None), (40, 4, 20, 56, None),
                                                        (20, 2, 10, 40, "simple"), (20, 2, 10, 40, "PH")])
def test_sim_imps_vs_oracle(sim_backend, Ls, cell, cut, target, spinful):
    Cs, Cl, _ = _case(Ls, cell, cut, {"chi_max": 8}, spinful)
    tp = {"chi_max": clean_chi(Cs, Cl, cell, cut, target, spinful)}
    Cs, Cl, ref = _case(Ls, cell, cut, tp, spinful)
    mps, err = slater.C_to_iMPS(Cs, Cl, tp, cell, cut, spinful=spinful, _backend=sim_backend, as_tenpy=False)
    assert mps.bc == "infinite" and mps.form == ["B"] * mps.L
    _compare(ref, mps, err)


def test_basis_rotation_matches_oracle():
    rng = np.random.default_rng(3)
    q = np.array([0, 0, 1, 1, 1, 2])
    Cm = rng.normal(size=(6, 6)) * (q[:, None] == q[None, :])
    S = np.sort(rng.uniform(size=6))[::-1]
    S /= np.linalg.norm(S)
    R1, u1, s1 = so.basis_rotation(Cm * 0.3, q, q, S, S)
    R2, u2, s2 = iMPS.basis_rotation(Cm * 0.3, S, S, "left", unitary_tol=10, schmidt_tol=10, q_bra=q, q_ket=q)
    assert np.allclose(R1, R2) and abs(u1 - u2) < 1e-14 and abs(s1 - s2) < 1e-14
    assert np.allclose(R2 @ R2.T, np.eye(6))


@pytest.mark.gpu
@pytest.mark.parametrize("Ls,cell,cut,tp", [(64, 2, 32, {"chi_max": 200}), (128, 2, 64, {"chi_max": 200}),
                                            (96, 4, 48, {"chi_max": 200, "svd_min": 1e-5}),
                                            # BASELINE configs[4], iMPS half: L_short = 1024, L_long = 1026, cut = 512
                                            (1024, 2, 512, {"chi_max": 1024, "svd_min": 1e-7})])
def test_gpu_imps_vs_oracle(gpu_backend, Ls, cell, cut, tp):
    """chi_max does not bind here (the gapped chain has ~40 Schmidt values above svd_min): the cut is set by the
    dynamic-range rule, away from the heavily degenerate multiplets of this spectrum."""
    Cs, Cl, ref = _case(Ls, cell, cut, tp)
    mps, err = slater.C_to_iMPS(Cs, Cl, tp, cell, cut, _backend=gpu_backend, as_tenpy=False)
    # unitary_error^2 = sum S^2 - sum |C S|^2 is a difference of two sums of O(1): with 512-site blocks the overlaps
    # of the two chains' Schmidt bases carry ~1e-12 relative rounding, which shows as ~1e-11 in the square (the
    # diagnostic reads 3.4e-6 where the reference's own truncation floor is 3.0e-7); the cell itself agrees to 1e-14
    fid = _compare(ref, mps, err, err2_tol=1e-13 if Ls <= 128 else 5e-11)
    print("iMPS cell fidelity", fid, "errors", err)


def _procrustes_vs_host(be, seed):
    """tmf_procrustes_blocks (block SVD kernel + U Vh + metric sums) against iMPS.basis_rotation on the same blocks."""
    import ctypes as C
    from temfpy_b200 import _lib
    rng = np.random.default_rng(seed)
    sizes = [(1, 1), (5, 5), (17, 17), (40, 40), (12, 9), (9, 12), (64, 64)]
    q_b = np.concatenate([np.full(m, i) for i, (m, _) in enumerate(sizes)])
    q_k = np.concatenate([np.full(n, i) for i, (_, n) in enumerate(sizes)])
    nb, nk = len(q_b), len(q_k)
    Cov = np.zeros((nb, nk))
    r0 = c0 = 0
    for m, n in sizes:        # near-isometric blocks (as in the iMPS conversion) with a little noise
        Q = np.linalg.qr(rng.normal(size=(max(m, n), max(m, n))))[0][:m, :n]
        Cov[r0: r0 + m, c0: c0 + n] = Q + 1e-3 * rng.normal(size=(m, n))
        r0, c0 = r0 + m, c0 + n
    Sk = np.sort(rng.uniform(0.01, 1.0, size=nk))[::-1]
    Sb = np.sort(rng.uniform(0.01, 1.0, size=nb))[::-1]
    R_ref, u_ref, s_ref = iMPS.basis_rotation(Cov, Sb, Sk, "left", unitary_tol=10, schmidt_tol=10, q_bra=q_b, q_ket=q_k)
    Cd, Skd = be.from_host(Cov.ravel()), be.from_host(Sk)
    Rd = be.from_host(np.zeros(nb * nk))
    met = be.from_host(np.zeros(2 * len(sizes)))
    jobs = (_lib.ProcrustesJob * len(sizes))()
    r0 = c0 = 0
    for u, (m, n) in enumerate(sizes):
        j = jobs[u]
        j.C = be.ptr(Cd) + 8 * (r0 * nk + c0)
        j.sk = be.ptr(Skd) + 8 * c0
        j.R = be.ptr(Rd) + 8 * (r0 * nk + c0)
        j.metrics = be.ptr(met) + 16 * u
        j.ldc = j.ldr = nk
        j.m, j.n = m, n
        r0, c0 = r0 + m, c0 + n
    wb = int(be.lib.tmf_procrustes_workspace(jobs, len(sizes)))
    work = be.empty(wb, np.uint8)
    _lib.check(be.lib, be.lib.tmf_procrustes_blocks(jobs, len(sizes), be.ptr(work), wb, be.stream))
    be.sync()
    R = be.to_host(Rd, nb * nk).reshape(nb, nk)
    mt = be.to_host(met, 2 * len(sizes)).reshape(-1, 2)
    assert np.abs(R - R_ref).max() < 1e-11, np.abs(R - R_ref).max()
    err2 = float(np.sum(Sk ** 2) - mt[:, 0].sum())
    assert abs(err2 - u_ref ** 2) < 1e-12 and abs(np.sqrt(mt[:, 1].sum()) - s_ref) < 1e-12


def test_sim_procrustes_blocks(sim_backend):
    _procrustes_vs_host(sim_backend, 3)


@pytest.mark.gpu
def test_gpu_procrustes_blocks(gpu_backend):
    _procrustes_vs_host(gpu_backend, 4)
