"""Pfaffian (Bogoliubov) path: oracle pinned against the reference's own outputs and exact known answers,
the kernels / driver executed by the CPU simulator (kernel logic), and the -m gpu parity tests through
the C ABI on the device."""
import ctypes as C

import numpy as np
import pytest

import pfaffian_oracle as po
import slater_oracle as so
from temfpy_b200 import _lib, pfaffian as pf
from tests import helpers

PF_FIXTURES = ["pfaffian_random_L8", "pfaffian_random_L9_real", "pfaffian_random_L14_chi20", "pfaffian_kitaev_L16"]


def _half_bonds(C_, tp, basis="C"):
    """bonds with eigenvalue-1/2 modes (vacuum parity = gauge there)."""
    L = len(C_) // 2
    out = set()
    for x in range(L + 1):
        m = po.bond_modes(C_, x, tp, basis, "L" if x <= L // 2 else "R", 0)
        if m.e.size and np.any(np.abs(m.e - 0.5) <= 1e-12):
            out.add(x)
    return out


# ---------------------------------------------------------------------------------------------
# oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", PF_FIXTURES)
def test_oracle_reproduces_reference_fixture(name):
    """The NumPy restatement equals the reference's own code on the same inputs."""
    g = helpers.golden(name)
    tp = helpers.golden_trunc(g)
    L, oc, basis = int(g["L"]), int(g["oc"]), str(g["basis"])
    Cm = po.correlation_matrix(g["H"], f"{basis}->{basis}")
    assert np.array_equal(Cm, g["C"])
    centre = po.bond_vectors_from_C(Cm, oc, tp, basis)
    parity = centre.modes.parity()
    vec = {oc: centre}
    for x in range(oc + 1, L + 1):
        vec[x] = po.bond_vectors_from_C(Cm, x, tp, basis, "R", parity)
    for x in range(oc):
        vec[x] = po.bond_vectors_from_C(Cm, x, tp, basis, "L", parity)
    for x in range(L + 1):
        v = vec[x]
        assert np.array_equal(v.modes.e, g[f"bond{x}_e"])
        assert np.array_equal(v.lam, g[f"bond{x}_lam"])
        assert np.array_equal(v.sets, g[f"bond{x}_sets"])
        assert v.modes.pL == int(g[f"bond{x}_pL"]) and v.modes.pR == int(g[f"bond{x}_pR"])
        assert np.array_equal(v.charges(v.modes.pL), g[f"bond{x}_charge"])
    for i in range(L):
        td = po.tensor_data(vec[i + 1], vec[i], "right") if i >= oc else po.tensor_data(vec[i], vec[i + 1], "left")
        assert np.allclose(td.N, g[f"site{i}_N"], rtol=0, atol=1e-13)
        assert abs(td.norm - float(g[f"site{i}_norm"])) < 1e-14
        assert td.qtotal == int(g[f"site{i}_qtotal"])
        assert np.allclose(po.dense_tensor(td), g[f"site{i}_T"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("L,seed,cplx", [(6, 1, True), (7, 2, True), (8, 3, False)])
def test_oracle_exact_correlators(L, seed, cplx):
    """Known answer: the Jordan-Wigner correlators of the dense MPS state reproduce the input Nambu
    correlation matrix (the check of the reference's examples/pfaffian.py:28-39)."""
    Cm = po.correlation_matrix(po.random_bdg(L, seed, cplx=cplx), "C->C")
    psi = so.mps_to_state(po.C_to_MPS(Cm, {"chi_max": 1000, "svd_min": 1e-7}, "C"))
    assert abs(np.linalg.norm(psi) - 1) < 1e-13
    G, F = po.state_correlators(psi)
    assert np.abs(G.T - Cm[::2, ::2]).max() < 5e-14        # C[2i, 2j]   = <c_j^+ c_i>  (pfaffian.py:14-25)
    assert np.abs(F.T - Cm[::2, 1::2]).max() < 5e-14       # C[2i, 2j+1] = <c_j c_i>


def test_oracle_pfaffian_known_answers():
    rng = np.random.default_rng(0)
    a, b, c, d, e, f = rng.normal(size=6) + 1j * rng.normal(size=6)
    A2 = np.array([[0, a], [-a, 0]])
    A4 = np.array([[0, a, b, c], [-a, 0, d, e], [-b, -d, 0, f], [-c, -e, -f, 0]])
    assert abs(po.pfaffian(A2) - a) < 1e-15
    assert abs(po.pfaffian(A4) - (a * f - b * e + c * d)) < 1e-14
    A = rng.normal(size=(7, 10, 10)) + 1j * rng.normal(size=(7, 10, 10))
    A = A - np.transpose(A, (0, 2, 1))
    assert np.abs(po.pfaffian(A) ** 2 - np.linalg.det(A)).max() < 1e-9 * np.abs(np.linalg.det(A)).max()
    assert po.pfaffian(np.zeros((0, 0))) == 1 and po.pfaffian(np.zeros((3, 3))) == 0


# ---------------------------------------------------------------------------------------------
# kernels + driver on the CPU simulator (kernel logic; same C ABI, same Python driver)
# ---------------------------------------------------------------------------------------------
def _pfaffians_through_abi(be, N, bra_masks, ket_masks, n1, n2, scale):
    lib = be.lib
    m = len(N)
    Nf = np.empty(2 * m * m)
    Nf[0::2], Nf[1::2] = N.real.ravel(), N.imag.ravel()
    Nd = be.from_host(Nf)
    bm = be.from_host(np.asarray(bra_masks, dtype=np.uint64).view(np.int64))
    km = be.from_host(np.asarray(ket_masks, dtype=np.uint64).view(np.int64))
    out = be.empty(2 * len(bra_masks) * len(ket_masks), np.float64)
    blk = (_lib.PfBlock * 1)()
    blk[0].N, blk[0].bra_masks, blk[0].ket_masks, blk[0].out = be.ptr(Nd), be.ptr(bm), be.ptr(km), be.ptr(out)
    blk[0].scale = scale
    blk[0].m, blk[0].n_bra, blk[0].n_ket, blk[0].n1, blk[0].n2 = m, len(bra_masks), len(ket_masks), n1, n2
    desc = be.empty(int(lib.tmf_pf_desc_bytes(1)), np.uint8)
    _lib.check(lib, lib.tmf_pfaffians_blocks(blk, 1, be.ptr(desc), be.stream))
    be.sync()
    o = be.to_host(out, 2 * len(bra_masks) * len(ket_masks))
    return (o[0::2] + 1j * o[1::2]).reshape(len(bra_masks), len(ket_masks))


def _random_block(rng, a1, a2, n1, n2, nb, nk):
    """random antisymmetric N over a2 ket + a1 bra modes and random masks with n1 / n2 excitations."""
    m = a1 + a2
    N = rng.normal(size=(m, m)) + 1j * rng.normal(size=(m, m))
    N = N - N.T

    def masks(count, lo, width, pop):
        out = []
        for _ in range(count):
            bits = rng.choice(width, size=pop, replace=False)
            out.append(sum(1 << int(lo + b) for b in bits))
        return np.array(out, dtype=np.uint64)
    return N, masks(nb, a2, a1, n1), masks(nk, 0, a2, n2)


def _check_pf_kernel(be, cases, seed):
    rng = np.random.default_rng(seed)
    for (a1, a2, n1, n2, nb, nk) in cases:
        N, bm, km = _random_block(rng, a1, a2, n1, n2, nb, nk)
        got = _pfaffians_through_abi(be, N, bm, km, n1, n2, 0.75)
        bits = lambda x: [t for t in range(a1 + a2) if (int(x) >> t) & 1]
        want = np.empty((nb, nk), dtype=complex)
        for i, b in enumerate(bm):
            for j, k in enumerate(km):
                idx = bits(int(k) | int(b))
                want[i, j] = 0.75 * po.pfaffian(N[np.ix_(idx, idx)])
        scale = np.abs(want).max() if want.size else 1.0
        assert np.abs(got - want).max() <= 1e-12 * max(scale, 1e-300), (a1, a2, n1, n2)


def test_sim_pfaffians_kernel(sim_backend):
    _check_pf_kernel(sim_backend, [(4, 4, 0, 0, 3, 5), (5, 6, 1, 1, 7, 9), (6, 5, 2, 4, 20, 33), (8, 8, 5, 5, 11, 40),
                                   (13, 12, 4, 6, 5, 130), (3, 3, 1, 2, 2, 2)], seed=5)


def test_sim_correlation_matrix(sim_backend):
    H = po.random_bdg(9, 11)
    for basis in ("C->C", "C->M", "M->M"):
        Hin = pf.matrix_C2M(H) if basis[0] == "M" else H
        assert np.abs(pf.correlation_matrix(Hin, basis, _backend=sim_backend) - po.correlation_matrix(Hin, basis)).max() < 1e-14


@pytest.mark.parametrize("name", PF_FIXTURES)
def test_sim_chain_vs_reference_fixture(sim_backend, name):
    g = helpers.golden(name)
    tp = helpers.golden_trunc(g)
    got = pf.C_to_MPS(g["C"], tp, basis=str(g["basis"]), _backend=sim_backend, as_tenpy=False)
    half = {x for x in range(int(g["L"]) + 1) if g[f"bond{x}_e"].size and np.any(np.abs(g[f"bond{x}_e"] - 0.5) <= 1e-12)}
    helpers.compare_pf_mps(helpers.golden_pf_mps(g), helpers.block_mps_to_dense(got), half)


def test_sim_chain_exact_correlators(sim_backend):
    Cm = po.correlation_matrix(po.random_bdg(9, 21), "C->C")
    m = pf.C_to_MPS(Cm, {"chi_max": 4096, "svd_min": 1e-7}, basis="C", _backend=sim_backend, as_tenpy=False)
    psi = so.mps_to_state(helpers.block_mps_to_dense(m))
    G, F = po.state_correlators(psi)
    assert np.abs(G.T - Cm[::2, ::2]).max() < 1e-13 and np.abs(F.T - Cm[::2, 1::2]).max() < 1e-13


@pytest.mark.parametrize("H,tp,basis", [
    (po.bdg_chain(24), {"chi_max": 32}, "C"),                      # eigenvalue-1/2 modes on odd bonds
    (po.bdg_chain(40, mu=0.3), {"chi_max": 48}, "C"),              # blocks > 64 rows: sketch / Ritz / Cholesky path
    (pf.matrix_C2M(po.random_bdg(12, 8)), {"chi_max": 40}, "M"),   # Majorana-basis input
])
def test_sim_chain_vs_oracle(sim_backend, H, tp, basis):
    Cm = po.correlation_matrix(H, f"{basis}->{basis}")
    ref = po.C_to_MPS(Cm, tp, basis)
    got = pf.H_to_MPS(H, tp, basis=basis, _backend=sim_backend, as_tenpy=False)
    helpers.compare_pf_mps(ref, helpers.block_mps_to_dense(got), _half_bonds(Cm, tp, basis))


def _finish_device_vs_host(be, monkeypatch):
    """tmf_pfaffian_site_finish (one CTA per site: Jacobi SVD of U*, inverse, N) against the same algebra in NumPy:
    same tensors."""
    for H, tp, oc in ((po.random_bdg(12, 8), {"chi_max": 40}, None), (po.bdg_chain(24), {"chi_max": 32}, 11)):
        Cm = po.correlation_matrix(H, "C->C")
        a = pf.C_to_MPS(Cm, tp, basis="C", ortho_center=oc, _backend=be, as_tenpy=False)
        monkeypatch.setenv("TMF_PF_HOST_FINISH", "1")
        b = pf.C_to_MPS(Cm, tp, basis="C", ortho_center=oc, _backend=be, as_tenpy=False)
        monkeypatch.delenv("TMF_PF_HOST_FINISH")
        assert a.meta["total_parity"] == b.meta["total_parity"]
        for i in range(a.L):
            Ta, Tb = a.get_B_dense(i), b.get_B_dense(i)
            assert Ta.shape == Tb.shape and np.abs(Ta - Tb).max() <= 1e-11 * max(np.abs(Tb).max(), 1e-300), i


def test_sim_site_finish_device_vs_host(sim_backend, monkeypatch):
    _finish_device_vs_host(sim_backend, monkeypatch)


def _site_norms_vs_oracle(be):
    """Onishi norms of the site tensors (pfaffian.py:1352-1359) against the oracle for a chain whose weakest
    entangled mode sits just above the cutoff (e = 3e-12): the non-entangled basis must be the exact complement of the
    *final* entangled columns (_PfChain._align_filled), otherwise the norms are off by theta^2 ~ 5e-8."""
    from temfpy_b200.schmidt_utils import to_stopping_condition
    L, cut, tp = 22, 10, {"chi_max": 32}
    Cl = po.correlation_matrix(po.bdg_chain(L, mu=0.3, delta=0.4), "C->C")
    trunc, t = po.Trunc.make(tp), to_stopping_condition(tp)
    chain = pf._PfChain(be, pf._prepare_CM(Cl, "C", t.svd_min ** 2), t, cut).run()
    centre = po.bond_vectors_from_C(Cl, cut, trunc, "C")
    assert centre.modes.e.min() < 1e-11
    par, prev = centre.modes.parity(), centre
    for i in range(cut, L):
        new = po.bond_vectors_from_C(Cl, i + 1, trunc, "C", "R", par)
        assert abs(po.tensor_data(new, prev, "right").norm - chain.site_tensors[i].norm) < 1e-12, i
        prev = new


def test_sim_site_norms_vs_oracle(sim_backend):
    _site_norms_vs_oracle(sim_backend)


def _centre_half_modes(be, L, oc=None):
    """Eigenvalue-1/2 Schmidt modes on the *central* bond (pfaffian.py:857-865): the symmetric chain cut at an
    odd bond.  The two real bases of the 1/2 space are paired by an SVD; the result equals the reference's state."""
    H = po.bdg_chain(L)
    Cm = po.correlation_matrix(H, "C->C")
    tp = {"chi_max": 32}
    x = oc or L // 2
    assert x in _half_bonds(Cm, tp), "the test case must carry 1/2 modes on its central bond"
    ref = po.C_to_MPS(Cm, tp, "C", ortho_center=oc)
    got = pf.C_to_MPS(Cm, tp, basis="C", ortho_center=oc, _backend=be, as_tenpy=False)
    return helpers.compare_pf_mps(ref, helpers.block_mps_to_dense(got), _half_bonds(Cm, tp))


@pytest.mark.parametrize("L,oc", [(10, None), (26, None), (24, 11)])
def test_sim_centre_half_modes(sim_backend, L, oc):
    _centre_half_modes(sim_backend, L, oc)


def test_sim_ortho_center_and_errors(sim_backend):
    Cm = po.correlation_matrix(po.random_bdg(10, 31), "C->C")
    tp = {"chi_max": 64}
    got = pf.C_to_MPS(Cm, tp, basis="C", ortho_center=3, _backend=sim_backend, as_tenpy=False)
    helpers.compare_pf_mps(po.C_to_MPS(Cm, tp, "C", ortho_center=3), helpers.block_mps_to_dense(got))
    with pytest.raises(ValueError):
        pf.C_to_MPS(Cm, tp, basis="X", _backend=sim_backend)
    with pytest.raises(ValueError):
        pf.C_to_MPS(Cm, tp, basis="C", unit_cell_width=3, _backend=sim_backend)
    with pytest.raises(ValueError, match="Bogoliubov vacuum"):          # a mixed state: C^2 != C
        pf.C_to_MPS(0.9 * Cm + 0.05 * np.eye(len(Cm)), tp, basis="C", _backend=sim_backend)


# ---------------------------------------------------------------------------------------------
# GPU parity tests (through the C ABI of the CUDA build)
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_pfaffians_kernel(gpu_backend):
    _check_pf_kernel(gpu_backend, [(4, 4, 0, 0, 3, 5), (5, 6, 1, 1, 70, 90), (6, 5, 2, 4, 200, 333),
                                   (8, 8, 5, 5, 110, 400), (13, 12, 4, 6, 50, 130), (16, 16, 8, 8, 40, 64)], seed=6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", PF_FIXTURES)
def test_gpu_chain_vs_reference_fixture(gpu_backend, name):
    g = helpers.golden(name)
    tp = helpers.golden_trunc(g)
    got = pf.C_to_MPS(g["C"], tp, basis=str(g["basis"]), _backend=gpu_backend, as_tenpy=False)
    half = {x for x in range(int(g["L"]) + 1) if g[f"bond{x}_e"].size and np.any(np.abs(g[f"bond{x}_e"] - 0.5) <= 1e-12)}
    helpers.compare_pf_mps(helpers.golden_pf_mps(g), helpers.block_mps_to_dense(got), half)


@pytest.mark.gpu
@pytest.mark.parametrize("H,tp", [
    (po.random_bdg(11, 41), {"chi_max": 4096, "svd_min": 1e-7}),
    (po.bdg_chain(48, mu=0.5, delta=0.2), {"chi_max": 64}),
    (po.bdg_chain(64, mu=0.0, delta=0.05), {"chi_max": 64}),
])
def test_gpu_chain_vs_oracle(gpu_backend, H, tp):
    Cm = po.correlation_matrix(H, "C->C")
    ref = po.C_to_MPS(Cm, tp, "C")
    got = pf.H_to_MPS(H, tp, basis="C", _backend=gpu_backend, as_tenpy=False)
    helpers.compare_pf_mps(ref, helpers.block_mps_to_dense(got), _half_bonds(Cm, tp))


@pytest.mark.gpu
def test_gpu_site_finish_device_vs_host(gpu_backend, monkeypatch):
    _finish_device_vs_host(gpu_backend, monkeypatch)
    _site_norms_vs_oracle(gpu_backend)


@pytest.mark.gpu
@pytest.mark.parametrize("L,oc", [(26, None), (50, None), (40, 19)])
def test_gpu_centre_half_modes(gpu_backend, L, oc):
    _centre_half_modes(gpu_backend, L, oc)


@pytest.mark.gpu
def test_gpu_cfg2_kitaev_chain_L128(gpu_backend):
    """BASELINE configs[1]: BdG p-wave chain L=128, t=1, mu=0, Delta=0.05, chi_max=128."""
    H = po.bdg_chain(128, t=1.0, mu=0.0, delta=0.05)
    tp = {"chi_max": 128}
    Cm = po.correlation_matrix(H, "C->C")
    ref = po.C_to_MPS(Cm, tp, "C")
    got = pf.H_to_MPS(H, tp, basis="C", _backend=gpu_backend, as_tenpy=False)
    rep = helpers.compare_pf_mps(ref, helpers.block_mps_to_dense(got), _half_bonds(Cm, tp))
    assert max(len(l) for l in got.lams) == 128
    print("cfg2 parity:", rep)


# ---------------------------------------------------------------------------------------------
# iMPS (pfaffian.py:1924-2242)
# ---------------------------------------------------------------------------------------------
def _imps_vs_oracle(be, Ls, cell, cut, tp, mu=0.3, delta=0.4):
    """pfaffian.C_to_iMPS against the oracle restatement: bond dimensions, Schmidt values, parity charges, the two
    error metrics and the unit cell itself (dominant eigenvalue of the mixed transfer matrix)."""
    from tests.test_imps import cell_transfer_eig
    import warnings
    Cs = po.correlation_matrix(po.bdg_chain(Ls, mu=mu, delta=delta), "C->C")
    Cl = po.correlation_matrix(po.bdg_chain(Ls + cell, mu=mu, delta=delta), "C->C")
    ref = po.C_to_iMPS(Cs, Cl, tp, cell, cut, "C")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mps, err = pf.C_to_iMPS(Cs, Cl, tp, cell, cut, basis="C", _backend=be, as_tenpy=False)
    assert mps.bc == "infinite" and mps.L == cell and mps.conserve == "parity"
    assert [len(l) for l in mps.lams] == [len(l) for l in ref.lams]
    for a, b in zip(ref.lams, mps.lams):
        assert np.all(np.abs(a - b) <= 1e-12 * a + np.minimum(1e-13 / (2 * a), 1e-8)), np.abs(a - b).max()
    for a, b in zip(ref.charges, mps.charges):
        assert np.array_equal(np.sort(a), np.sort(b))
    got = [mps.get_B_dense(i) for i in range(mps.L)]
    for i, T in enumerate(got):         # parity rule of every tensor: q(vL) + p - qtotal = q(vR) (mod 2)
        a, p, b = np.nonzero(np.abs(T) > 1e-12)
        d = np.unique((np.asarray(mps.charges[i])[a] + p - np.asarray(mps.charges[i + 1])[b]) % 2)
        assert d.size == 1, (i, d)
    # (unitary_error^2 = sum S^2 - sum |C S|^2 is a difference of two O(1) sums: agreement at the 1e-11 level)
    assert abs(err.left_unitary ** 2 - ref.errors[0] ** 2) < 3e-11 and abs(err.left_schmidt - ref.errors[1]) < 1e-9
    e_mix = cell_transfer_eig(ref.tensors, got)
    fid = abs(e_mix) ** 2 / abs(cell_transfer_eig(ref.tensors, ref.tensors) * cell_transfer_eig(got, got))
    assert fid >= 1 - 1e-9, fid
    return fid, err


@pytest.mark.parametrize("Ls,cell,cut,tp", [(20, 2, 10, {"chi_max": 32}), (18, 1, 9, {"chi_max": 24}),
                                            (24, 4, 12, {"chi_max": 40})])
def test_sim_imps_vs_oracle(sim_backend, Ls, cell, cut, tp):
    _imps_vs_oracle(sim_backend, Ls, cell, cut, tp)


@pytest.mark.gpu
@pytest.mark.parametrize("Ls,cell,cut,tp,mu", [(32, 2, 16, {"chi_max": 64}, 0.3), (64, 2, 32, {"chi_max": 96}, 2.5),
                                               (96, 3, 45, {"chi_max": 64}, -2.3)])
def test_gpu_imps_vs_oracle(gpu_backend, Ls, cell, cut, tp, mu):
    # (|mu| > 2 t: the trivial phase -- the topological one has Majorana zero modes, split by ~exp(-L), which
    #  correlation_matrix refuses at these lengths like the reference, pfaffian.py:377-380)
    print(_imps_vs_oracle(gpu_backend, Ls, cell, cut, tp, mu=mu))


def test_oracle_imps_known_answers():
    """The oracle restatement of pfaffian.C_to_iMPS (pfaffian.py:1924-2091; the reference driver itself needs TeNPy):
    the unit cell reproduces the bulk density of the long chain, is unitarily gauged and has a normalised transfer
    matrix."""
    import slater_oracle as so
    from tests.test_imps import cell_transfer_eig
    L, cell, cut = 24, 2, 12
    Cs = po.correlation_matrix(po.bdg_chain(L, mu=2.5, delta=0.4), "C->C")
    Cl = po.correlation_matrix(po.bdg_chain(L + cell, mu=2.5, delta=0.4), "C->C")
    im = po.C_to_iMPS(Cs, Cl, {"chi_max": 64}, cell, cut, "C")
    assert im.errors[0] < 1e-6 and im.errors[1] < 1e-9
    n, I = np.diag([0, 1.0]), np.eye(2)
    assert abs(so.imps_expectation(im, [n, I]).real - Cl[2 * cut, 2 * cut].real) < 1e-6
    assert abs(so.imps_expectation(im, [I, n]).real - Cl[2 * cut + 2, 2 * cut + 2].real) < 1e-6
    assert abs(abs(cell_transfer_eig(im.tensors, im.tensors)) - 1) < 1e-6
