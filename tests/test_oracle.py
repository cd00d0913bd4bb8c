"""The oracle restatement against (1) fixtures produced by the reference's own code and
(2) exact known answers.  CPU only."""
import numpy as np
import pytest

import slater_oracle as so
from tests import helpers

CASES = ["slater_random_L12", "slater_random_L20_chi24", "slater_random_L11_N4", "slater_chain_L16",
         "slater_random_L40", "slater_complex_L12", "slater_complex_L24_chi32"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    g = helpers.golden(name)
    tp = helpers.golden_trunc(g)
    C, N = so.correlation_matrix(g["H"], int(g["N"]))
    assert np.array_equal(C, g["C"])
    L, oc = int(g["L"]), int(g["oc"])
    trunc = so.Trunc.make(tp)
    # real inputs: bit for bit (same NumPy calls in the same order); complex inputs: the spectra are, the site
    # quantities agree to a few ulp (the reference's HT(v) @ w and our v.conj().T @ w take different BLAS paths)
    same = (lambda a, b: np.allclose(a, b, rtol=0, atol=1e-13)) if np.iscomplexobj(C) else np.array_equal
    centre = so.bond_vectors_from_C(C, oc, trunc, "LR")
    prev, bonds = centre, {oc: centre}
    for i in range(oc, L):
        new = so.bond_vectors_from_C(C, i + 1, trunc, "R")
        td = so.tensor_data(new, prev, "right")
        assert same(td.S, g[f"site{i}_S"]) and same(td.det_always, g[f"site{i}_det"])
        assert np.array_equal(td.sets_bra, g[f"site{i}_sets_bra"])
        assert np.array_equal(td.sets_ket, g[f"site{i}_sets_ket"])
        bonds[i + 1] = prev = new
    prev = centre
    for i in reversed(range(oc)):
        new = so.bond_vectors_from_C(C, i, trunc, "L")
        td = so.tensor_data(new, prev, "left")
        assert same(td.S, g[f"site{i}_S"]) and same(td.det_always, g[f"site{i}_det"])
        assert np.array_equal(td.sets_bra, g[f"site{i}_sets_bra"])
        bonds[i] = prev = new
    for x, v in bonds.items():      # bit-exact: same NumPy calls in the same order
        assert np.array_equal(v.modes.e, g[f"bond{x}_e"])
        assert np.array_equal(v.lam, g[f"bond{x}_lam"])
        assert np.array_equal(v.n_left, g[f"bond{x}_charge"])
    ref = helpers.golden_dense_mps(g)
    mine = so.C_to_MPS(C, tp)
    for a, b in zip(ref.tensors, mine.tensors):
        assert same(a, b)


def test_lowest_sums_fixture():
    g = helpers.golden("lowest_sums")
    for c in range(int(g["ncases"])):
        chi, svd_min, fl, fr = g[f"c{c}_par"]
        sec = [int(s) for s in g[f"c{c}_sectors"]] if bool(g[f"c{c}_has_sectors"]) else None
        tp = so.Trunc(sectors=sec, chi_max=None if chi < 0 else int(chi), svd_min=float(svd_min))
        sums, sets = so.lowest_sums(g[f"c{c}_a"], tp, None if fl < 0 else int(fl), None if fr < 0 else int(fr))
        assert np.array_equal(sums, g[f"c{c}_sums"])
        assert np.array_equal(sets.reshape(g[f"c{c}_sets"].shape), g[f"c{c}_sets"])


@pytest.mark.parametrize("L,N,seed", [(8, None, 8), (10, None, 10), (11, 4, 11), (12, 7, 12)])
def test_exact_slater_amplitudes(L, N, seed):
    """Known answer: psi(occ) = det Phi[occ, :] (SURVEY 8c pin (i))."""
    H = helpers.random_hamiltonian(L, seed)
    C, n = so.correlation_matrix(H, N)
    Phi = np.linalg.eigh(H)[1][:, :n]
    mps = so.C_to_MPS(C, {"chi_max": 4096, "svd_min": 1e-7})
    psi = so.mps_to_state(mps)
    assert abs(abs(np.vdot(so.exact_slater_state(Phi), psi)) - 1) < 1e-13
    full = all(c == min(2 ** x, 2 ** (L - x)) for x, c in enumerate(mps.chi))
    for i, T in enumerate(mps.tensors if full else []):     # canonical-form residuals (untruncated)
        if i < mps.ortho_center:
            E = np.einsum("apb,apc->bc", T, T)
        else:
            E = np.einsum("apb,cpb->ac", T, T)
        assert np.abs(E - np.eye(len(E))).max() < 1e-13


def test_oracle_complex_hamiltonian():
    rng = np.random.default_rng(5)
    L = 9
    H = rng.normal(size=(L, L)) + 1j * rng.normal(size=(L, L))
    H = H + H.conj().T
    C, n = so.correlation_matrix(H, 5)
    Phi = np.linalg.eigh(H)[1][:, :n]
    psi = so.mps_to_state(so.C_to_MPS(C, {"chi_max": 4096, "svd_min": 1e-7}))
    assert abs(abs(np.vdot(so.exact_slater_state(Phi), psi)) - 1) < 1e-13


def test_live_reference_agreement():
    """When /root/reference is present (build container) compare with the live reference too."""
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference checkout not present")
    ref = ref_shim.load()
    C, _ = so.correlation_matrix(so.hopping_chain(64))
    tp = {"chi_max": 64}
    for x, which in [(32, "LR"), (33, "R"), (20, "L"), (64, "R"), (0, "L")]:
        r = ref.slater.SchmidtVectors.from_correlation_matrix(C, x, tp, which=which)
        o = so.bond_vectors_from_C(C, x, tp, which)
        assert np.array_equal(r.schmidt_values, o.lam)
        assert {int(k): (int(s.start), int(s.stop)) for k, s in r.idx_L.items()} == o.idx_L
