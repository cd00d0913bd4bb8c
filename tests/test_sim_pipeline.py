"""The CUDA kernel sources executed by the CPU simulator (tests/hostsim) through the same C ABI and
the same Python driver as on the GPU; compared with the oracle / the reference fixtures.  CPU only.
These tests check kernel *logic* (indexing, pivoting, signs, planning); the DMMA tile code itself
is covered by the -m gpu tests."""
import ctypes as C

import numpy as np
import pytest

import slater_oracle as so
from temfpy_b200 import _lib
from tests import helpers


@pytest.mark.parametrize("name", ["slater_random_L12", "slater_random_L20_chi24", "slater_random_L11_N4",
                                  "slater_random_L40"])
def test_chain_vs_reference_fixture(sim_backend, name):
    g = helpers.golden(name)
    tp = helpers.golden_trunc(g)
    res = helpers.run_native(sim_backend, g["C"], tp, int(g["N"]))
    rep = helpers.compare_mps(helpers.golden_dense_mps(g), helpers.chain_to_dense(res), tp)
    assert rep["ambiguous"] == []


def test_chain_exact_amplitudes(sim_backend):
    L = 10
    H = helpers.random_hamiltonian(L, 77)
    Cm, n = so.correlation_matrix(H)
    Phi = np.linalg.eigh(H)[1][:, :n]
    res = helpers.run_native(sim_backend, Cm, {"chi_max": 4096, "svd_min": 1e-7}, n)
    psi = so.mps_to_state(helpers.chain_to_dense(res))
    assert abs(abs(np.vdot(so.exact_slater_state(Phi), psi)) - 1) < 1e-13


def test_chain_sketch_path_vs_oracle(sim_backend):
    """Blocks larger than 64 sites go through the range-sketch / Rayleigh-Ritz / Cholesky path."""
    L = 150
    Cm, n = so.correlation_matrix(helpers.random_hamiltonian(L, 5))
    tp = {"chi_max": 48}
    res = helpers.run_native(sim_backend, Cm, tp, n)
    rep = helpers.compare_mps(so.C_to_MPS(Cm, tp), helpers.chain_to_dense(res), tp)
    assert rep["ambiguous"] == []


def test_chain_ortho_center_and_shards(sim_backend):
    """Sharded conversion == unsharded, bit for bit (multi-GPU without a cluster)."""
    L = 30
    Cm, n = so.correlation_matrix(helpers.random_hamiltonian(L, 9))
    tp = {"chi_max": 32}
    full = helpers.run_native(sim_backend, Cm, tp, n, ortho_center=11)
    ref = so.C_to_MPS(Cm, tp, ortho_center=11)
    helpers.compare_mps(ref, helpers.chain_to_dense(full), tp)
    for lo, hi in [(0, 7), (7, 19), (19, 30)]:
        part = helpers.run_native(sim_backend, Cm, tp, n, ortho_center=11, site_lo=lo, site_hi=hi)
        for x in range(lo, hi + 1):
            assert np.array_equal(part.bonds[x].schmidt_values, full.bonds[x].schmidt_values)
            assert np.array_equal(part.bonds[x].charge, full.bonds[x].charge)
        for i in range(lo, hi):
            # tensors may differ by the sign gauge of a boundary bond that only one shard pairs
            a, b = part.sites[i].dense(), full.sites[i].dense()
            assert np.allclose(np.abs(a), np.abs(b), atol=1e-13)


def test_minors_kernel_vs_tensor_block(sim_backend):
    """K10 stage test: same S / masks as the oracle -> every block equal to _tensor_block."""
    lib = sim_backend.lib
    Cm, _ = so.correlation_matrix(helpers.random_hamiltonian(48, 4))
    tp = {"chi_max": 64}
    ket = so.bond_vectors_from_C(Cm, 24, tp)
    bra = so.bond_vectors_from_C(Cm, 25, tp, "R")
    td = so.tensor_data(bra, ket, "right")
    sb, sk = td.S.shape
    S = np.asfortranarray(td.S).ravel(order="F").copy()
    pack = lambda sets: (sets.astype(np.uint64) << np.arange(sets.shape[1], dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint64)
    bm, km = pack(td.sets_bra), pack(td.sets_ket)
    det = np.array([td.det_always])
    blocks, outs, want = [], [], []
    for q in np.unique(td.q_ket):
        kr = np.flatnonzero(td.q_ket == q); br = np.flatnonzero(td.q_bra == q - td.qtotal)
        if not br.size:
            continue
        out = np.zeros(br.size * kr.size)
        b = _lib.MinorBlock(S=S.ctypes.data, det=det.ctypes.data, bra_masks=bm[br[0]:].ctypes.data,
                            ket_masks=km[kr[0]:].ctypes.data, out=out.ctypes.data, s_bra=sb, s_ket=sk,
                            n_bra=br.size, n_ket=kr.size, minor=int(td.sets_ket[kr[0]].sum()))
        blocks.append(b); outs.append(out)
        want.append(td.det_always * so.tensor_block(td.S, td.sets_bra[br], td.sets_ket[kr]))
    arr = (_lib.MinorBlock * len(blocks))(*blocks)
    desc = np.zeros(lib.tmf_minor_desc_bytes(len(blocks)), np.uint8)
    _lib.check(lib, lib.tmf_minors_blocks(arr, len(blocks), desc.ctypes.data, None))
    for o, w in zip(outs, want):
        assert np.allclose(o.reshape(w.shape), w, rtol=0, atol=1e-14)


def test_gemm_grouped_semantics(sim_backend):
    lib = sim_backend.lib
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.normal(size=(70, 33))); B = np.asfortranarray(rng.normal(size=(33, 45)))
    Cc = np.asfortranarray(rng.normal(size=(70, 45))); C0 = Cc.copy()
    j = _lib.GemmJob(A=A.ctypes.data, B=B.ctypes.data, C=Cc.ctypes.data, M=70, N=45, K=33, lda=70, ldb=33,
                     ldc=70, transA=0, transB=0, alpha=2.0, beta=-1.0)
    desc = np.zeros(lib.tmf_gemm_desc_bytes(1), np.uint8)
    _lib.check(lib, lib.tmf_gemm_grouped((_lib.GemmJob * 1)(j), 1, desc.ctypes.data, None))
    assert np.allclose(Cc, 2 * A @ B - C0, atol=1e-12)


@pytest.mark.parametrize("tp", [{"chi_max": 64}, {"chi_max": 200, "svd_min": 1e-5}, {"chi_max": 7},
                                {"chi_max": 64, "sectors": [q for q in range(31) if q not in (14, 17)]}])
def test_device_enumeration_equals_host_enumeration(sim_backend, monkeypatch, tp):
    """The enumeration kernel (bucket queue, one warp per bond) reproduces the host implementation of
    schmidt_utils.lowest_sums / SchmidtVectors.from_schmidt_modes bit for bit: masks in the same order,
    identical Schmidt values, charges and sector tables."""
    L = 30
    Cm, n = so.correlation_matrix(helpers.random_hamiltonian(L, 21))
    dev = helpers.run_native(sim_backend, Cm, tp, n)
    monkeypatch.setenv("TMF_HOST_ENUMERATE", "1")
    host = helpers.run_native(sim_backend, Cm, tp, n)
    for x in range(L + 1):
        a, b = dev.bonds[x], host.bonds[x]
        assert np.array_equal(a.masks, b.masks)
        assert np.array_equal(a.schmidt_values, b.schmidt_values)
        assert np.array_equal(a.charge, b.charge)
        assert a.idx_L == b.idx_L


def test_decoupled_blocks_rank_deficient_sketch(sim_backend):
    """Two decoupled subsystems: at and near the cut between them the off-diagonal block of C has rank
    0 ... few, far below the sketch width -- the Cholesky-QR panels meet exactly dependent / zero columns and
    must hand over to the MGS2 body, the mode count drops to zero at the decoupled bond."""
    La, Lb = 80, 70
    H = np.zeros((La + Lb, La + Lb))
    H[:La, :La] = helpers.random_hamiltonian(La, 1)
    H[La:, La:] = helpers.random_hamiltonian(Lb, 2)
    Cm, n = so.correlation_matrix(H)
    tp = {"chi_max": 32}
    res = helpers.run_native(sim_backend, Cm, tp, n, ortho_center=70)
    assert res.bonds[La].k == 0 and res.bonds[La].chi == 1
    rep = helpers.compare_mps(so.C_to_MPS(Cm, tp, ortho_center=70), helpers.chain_to_dense(res), tp)
    assert rep["ambiguous"] == []


@pytest.mark.parametrize("case", ["anderson", "svd_min_1e-3"])
def test_structured_inputs_vs_oracle(sim_backend, case):
    """Regression tests of two robustness bugs: (1) localised (Anderson) states -- the singular values of the
    off-diagonal block fall below the rounding noise within one sketch panel; normalised noise columns used to
    re-enter Q and entangled modes were lost; (2) svd_min = 1e-3 -- near-empty modes (e < svd_min^2) stay in
    the filled-space projector and the rank tolerance of the pivoted Cholesky has to sit above them."""
    rng = np.random.default_rng(5)
    L = 130
    H = np.zeros((L, L))
    i = np.arange(L - 1)
    H[i, i + 1] = H[i + 1, i] = -1.0
    if case == "anderson":
        H += np.diag(2.0 * rng.standard_normal(L))
        tp = {"chi_max": 32}
    else:
        H += np.diag(0.1 * rng.standard_normal(L))
        tp = {"svd_min": 1e-3}
    Cm, n = so.correlation_matrix(H)
    res = helpers.run_native(sim_backend, Cm, tp, n)
    rep = helpers.compare_mps(so.C_to_MPS(Cm, tp), helpers.chain_to_dense(res), tp)
    # (with an svd_min-limited cut a near-degenerate multiplet may straddle the cut: compare_mps audits it)
    assert rep["ambiguous"] == [] or "svd_min" in case


@pytest.mark.parametrize("case", ["random", "chain", "oc"])
def test_device_site_plans_match_host(sim_backend, case):
    """The device planner (plan.cu) against the host planner (hostlogic.cpp site_plan, the restatement of
    slater.py:760-825, :1027-1058, :1106-1141): identical headers, block tables and tensors."""
    kw = {}
    if case == "random":
        Cm, n = so.correlation_matrix(helpers.random_hamiltonian(60, 4))
        tp = {"chi_max": 48}
    elif case == "chain":
        Cm, n = so.correlation_matrix(so.hopping_chain(40))
        tp = {"chi_max": 32, "svd_min": 1e-7}
    else:
        Cm, n = so.correlation_matrix(helpers.random_hamiltonian(50, 9, 3.0))
        tp, kw = {"chi_max": 40}, dict(ortho_center=13)
    dev = helpers.run_native(sim_backend, Cm, tp, n, device_plan=True, **kw)
    host = helpers.run_native(sim_backend, Cm, tp, n, device_plan=False, **kw)
    assert dev.stats["path"] == dict(nested=True, device_plan=True)
    assert host.stats["path"] == dict(nested=True, device_plan=False)
    for i in range(len(Cm)):
        a, b = dev.sites[i], host.sites[i]
        for f, _ in a.plan._fields_:
            assert getattr(a.plan, f) == getattr(b.plan, f), (i, f)
        assert np.array_equal(a.row_p, b.row_p) and np.array_equal(a.row_alpha, b.row_alpha)
        assert len(a.blocks) == len(b.blocks)
        for ba, bb in zip(a.blocks, b.blocks):
            assert ba[:5] == bb[:5] and np.array_equal(ba[5], bb[5])
    for x in range(len(Cm) + 1):
        assert np.array_equal(dev.bonds[x].masks, host.bonds[x].masks)
        assert np.array_equal(dev.bonds[x].schmidt_values, host.bonds[x].schmidt_values)
        assert dev.bonds[x].idx_L == host.bonds[x].idx_L


@pytest.mark.parametrize("name", ["slater_complex_L12", "slater_complex_L24_chi32"])
def test_complex_reference_fixtures(sim_backend, name):
    """Complex Hamiltonians (the reference's own example, examples/slater.py:15-27): embedded mode extraction,
    complex pairing, complex nested site stage and complex minors against fixtures of the live reference."""
    g = helpers.golden(name)
    tp = helpers.golden_trunc(g)
    res = helpers.run_native(sim_backend, g["C"], tp, int(g["N"]))
    assert res.sites[0].blocks[0][5].dtype == np.complex128
    helpers.compare_mps(helpers.golden_dense_mps(g), helpers.chain_to_dense(res), tp)


@pytest.mark.parametrize("L,N,seed", [(8, 5, 8), (11, None, 11)])
def test_complex_exact_amplitudes(sim_backend, L, N, seed):
    """Known answer for complex orbitals: psi(occ) = det Phi[occ, :] (SURVEY 8c pin (i): L = 8 complex N = 5, L = 11 complex)."""
    H = helpers.random_hamiltonian(L, seed, cplx=True)
    Cm, n = so.correlation_matrix(H, N)
    Phi = np.linalg.eigh(H)[1][:, :n]
    res = helpers.run_native(sim_backend, Cm, {"chi_max": 4096, "svd_min": 1e-7}, n)
    psi = so.mps_to_state(helpers.chain_to_dense(res))
    assert abs(abs(np.vdot(so.exact_slater_state(Phi), psi)) - 1) < 1e-13


def test_complex_public_api_vs_oracle(sim_backend):
    from temfpy_b200 import slater, testing
    old, testing.TEST_ACTION = testing.TEST_ACTION, "pass"
    try:
        H = helpers.random_hamiltonian(36, 2, cplx=True)
        C, N = slater.correlation_matrix(H, _backend=sim_backend)
        Co, No = so.correlation_matrix(H)
        assert N == No and np.abs(C - Co).max() < 1e-14 and np.iscomplexobj(C)
        tp = {"chi_max": 40}
        mps = slater.H_to_MPS(H, tp, as_tenpy=False, _backend=sim_backend, ortho_center=13)
        rep = helpers.compare_mps(so.C_to_MPS(Co, tp, ortho_center=13), helpers.block_mps_to_dense(mps), tp)
        assert rep["overlap"] >= 1 - 1e-10
    finally:
        testing.TEST_ACTION = old


def test_coarse_truncation_uses_exact_site_stage(sim_backend):
    """svd_min = 1e-3 (cutoff 1e-6): the nested-projector site stage would be off by O(cutoff) (overlap with the
    reference 1 - 3e-11 here, 1 - 4e-9 in the worst case tools/campaign_sim.py found); run_chain switches to the
    explicit filled bases and agrees to rounding."""
    from temfpy_b200 import engine
    from temfpy_b200.schmidt_utils import to_stopping_condition
    rng = np.random.default_rng(2)
    L = 100
    H = np.zeros((L, L))
    i = np.arange(L - 1)
    H[i, i + 1] = H[i + 1, i] = -1.0
    H += np.diag(2.0 * rng.standard_normal(L))
    tp = {"chi_max": 16, "svd_min": 1e-3}
    Cm, n = so.correlation_matrix(H)
    ref = so.C_to_MPS(Cm, tp)
    res = engine.run_chain(sim_backend, np.ascontiguousarray(Cm).ravel(), L, L, to_stopping_condition(tp), n)
    assert res.options["nested"] is False
    rep = helpers.compare_mps(ref, helpers.chain_to_dense(res), tp)
    assert abs(1 - rep["overlap"]) < 1e-13, rep
