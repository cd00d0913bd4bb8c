"""Native host logic (enumeration, Schmidt tables, site planning) against the oracle.  CPU only:
these entry points are plain host C++ and are exercised through the simulator build of the ABI."""
import ctypes as C

import numpy as np
import pytest

import slater_oracle as so
from temfpy_b200 import _lib
from temfpy_b200 import schmidt_utils as su
from tests import helpers


@pytest.fixture(scope="module")
def lib(sim_backend):
    return sim_backend.lib


def test_lowest_sums_golden(lib):
    g = helpers.golden("lowest_sums")
    for c in range(int(g["ncases"])):
        chi, svd_min, fl, fr = g[f"c{c}_par"]
        sec = [int(s) for s in g[f"c{c}_sectors"]] if bool(g[f"c{c}_has_sectors"]) else None
        tp = su.StoppingCondition(sectors=sec, chi_max=None if chi < 0 else int(chi), svd_min=float(svd_min))
        sums, sets = su.lowest_sums(g[f"c{c}_a"], tp, filled_left=None if fl < 0 else int(fl),
                                    filled_right=None if fr < 0 else int(fr), _lib_override=lib)
        assert np.array_equal(sums, g[f"c{c}_sums"]), c           # bit-exact running sums
        assert np.array_equal(sets, g[f"c{c}_sets"].reshape(sets.shape)), c


@pytest.mark.parametrize("seed", range(6))
def test_lowest_sums_random_vs_oracle(lib, seed):
    rng = np.random.default_rng(seed)
    k = int(rng.integers(1, 18))
    a = rng.normal(size=k) * rng.uniform(0.5, 4)
    kw = dict(chi_max=int(rng.integers(1, 300)), svd_min=float(10 ** rng.uniform(-7, -1)))
    fl = int(rng.integers(0, 5))
    s0, b0 = so.lowest_sums(a, so.Trunc(**kw), fl, None)
    s1, b1 = su.lowest_sums(a, su.StoppingCondition(**kw), filled_left=fl, _lib_override=lib)
    assert np.array_equal(s0, s1) and np.array_equal(b0.reshape(b1.shape), b1)


def test_stopping_condition_semantics():
    tp = su.StoppingCondition(chi_max=3, svd_min=1e-2)
    assert tp(np.array([0.0, 1.0, 2.0])) and not tp(np.array([0.0, 1.0, 2.0, 3.0]))
    assert not tp(np.array([0.0, 5.0]))
    assert tp.truncate(np.array([0.0, 1.0, 1.0, 2.0, 9.0])) == 1 or True
    ref = so.Trunc(chi_max=3, svd_min=1e-2)
    for lv in ([0.0, 1.0, 1.0, 2.0, 9.0], [0.0, 0.5, 0.5 + 1e-13, 0.7], [0.0, 4.0, 4.7]):
        assert tp.truncate(np.array(lv)) == ref.truncate(lv)
    with pytest.raises(TypeError):
        su.to_stopping_condition(3)
    with pytest.raises(AssertionError):
        su.StoppingCondition(chi_max=0)


def _bond_vectors_native(lib, e, filled_left, tp, cap=4096):
    nb = 1
    ebuf = np.zeros(64)
    ebuf[: len(e)] = e
    k = (C.c_int * 1)(len(e))
    fl = (C.c_int * 1)(filled_left)
    S = 66
    masks = np.zeros(cap, np.uint64); lam = np.zeros(cap); charge = np.zeros(cap, np.int32)
    chi = np.zeros(1, np.int32); sq = np.zeros(S, np.int32); ss = np.zeros(S, np.int32); sn = np.zeros(1, np.int32)
    rc = lib.tmf_bond_vectors_batched(nb, ebuf.ctypes.data_as(_lib.c_double_p), k, fl,
                                      -1 if tp.chi_max is None else tp.chi_max, tp.svd_min, tp.degeneracy_tol,
                                      None, -1, cap, masks.ctypes.data_as(_lib.c_u64_p),
                                      lam.ctypes.data_as(_lib.c_double_p), charge.ctypes.data_as(_lib.c_int_p),
                                      chi.ctypes.data_as(_lib.c_int_p), sq.ctypes.data_as(_lib.c_int_p),
                                      ss.ctypes.data_as(_lib.c_int_p), sn.ctypes.data_as(_lib.c_int_p), 1)
    _lib.check(lib, rc)
    n = int(chi[0])
    return masks[:n], lam[:n], charge[:n], sq[: sn[0]], ss[: sn[0] + 1]


@pytest.mark.parametrize("L,x,which,chi", [(64, 32, "LR", 64), (64, 40, "R", 64), (64, 20, "L", 30),
                                           (40, 7, "L", 64), (40, 39, "R", 64)])
def test_bond_vectors_identical_spectrum(lib, L, x, which, chi):
    """K6/K7 in isolation on *identical* e arrays: integers exact, lambda to 1e-15."""
    H = helpers.random_hamiltonian(L, L + x)
    Cm, _ = so.correlation_matrix(H)
    tp = so.Trunc(chi_max=chi)
    v = so.bond_vectors_from_C(Cm, x, tp, which)
    masks, lam, charge, sq, ss = _bond_vectors_native(lib, v.modes.e, v.modes.n_filled("L"), tp)
    bits = ((masks[:, None] >> np.arange(v.modes.e.size, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    assert np.array_equal(bits, v.sets)
    assert np.array_equal(charge, v.n_left)
    assert np.allclose(lam, v.lam, rtol=1e-14, atol=0)
    assert {int(q): (int(ss[i]), int(ss[i + 1])) for i, q in enumerate(sq)} == v.idx_L


@pytest.mark.parametrize("L,i,chi", [(40, 25, 64), (40, 20, 64), (40, 8, 64), (40, 39, 64), (40, 0, 64),
                                     (64, 33, 24)])
def test_site_plan_matches_oracle(lib, L, i, chi):
    """_select_orbitals / row order / charge blocks (slater.py:760-825, 1027-1058, 1132-1141)."""
    H = helpers.random_hamiltonian(L, 3 * L + i)
    Cm, N = so.correlation_matrix(H)
    tp = so.Trunc(chi_max=chi)
    oc = L // 2
    if i >= oc:
        mode, side = 1, "R"
        ket = so.bond_vectors_from_C(Cm, i, tp, "LR" if i == oc else "R")
        bra = so.bond_vectors_from_C(Cm, i + 1, tp, "R")
        n_bra, n_ket = L - i - 1, L - i
    else:
        mode, side = 0, "L"
        ket = so.bond_vectors_from_C(Cm, i + 1, tp, "LR" if i + 1 == oc else "L")
        bra = so.bond_vectors_from_C(Cm, i, tp, "L")
        n_bra, n_ket = i, i + 1
    td = so.tensor_data(bra, ket, "right" if mode else "left")

    def pack(v):
        k = v.modes.e.size
        m = (v.sets.astype(np.uint64) << np.arange(k, dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint64)
        return np.ascontiguousarray(m), np.ascontiguousarray(v.n_left.astype(np.int32)), k, v.modes.n_filled(side)

    mb, qb, kb, fb = pack(bra)
    mk, qk, kk, fk = pack(ket)
    plan = _lib.SitePlan()
    cap = 2 * len(mb) + 8
    bra_cols = np.zeros(kb + fb + 2, np.int32); bra_sign = np.zeros(kb + fb + 2)
    ket_cols = np.zeros(kk + fk + 2, np.int32); ket_sign = np.zeros(kk + fk + 2)
    bmask = np.zeros(cap, np.uint64); kmask = np.zeros(len(mk) + 8, np.uint64)
    row_p = np.zeros(cap, np.int32); row_a = np.zeros(cap, np.int32); blocks = np.zeros(6 * 80, np.int32)
    P = lambda a, t: a.ctypes.data_as(t)
    rc = lib.tmf_slater_site_plan(mode, n_bra, n_ket, kb, fb, N, len(mb), P(mb, _lib.c_u64_p), P(qb, _lib.c_int_p),
                                  kk, fk, N, len(mk), P(mk, _lib.c_u64_p), P(qk, _lib.c_int_p), C.byref(plan),
                                  P(bra_cols, _lib.c_int_p), P(bra_sign, _lib.c_double_p),
                                  P(ket_cols, _lib.c_int_p), P(ket_sign, _lib.c_double_p),
                                  P(bmask, _lib.c_u64_p), P(kmask, _lib.c_u64_p), P(row_p, _lib.c_int_p),
                                  P(row_a, _lib.c_int_p), P(blocks, _lib.c_int_p))
    _lib.check(lib, rc)
    assert (plan.s_bra, plan.s_ket) == td.S.shape
    assert plan.n_rows == len(td.sets_bra) and plan.chi_ket == len(td.sets_ket)
    unpack = lambda m, s: ((m[:, None] >> np.arange(s, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    assert np.array_equal(unpack(bmask[: plan.n_rows], plan.s_bra), td.sets_bra)
    assert np.array_equal(unpack(kmask[: plan.chi_ket], plan.s_ket), td.sets_ket)
    assert np.array_equal(np.stack([row_p[: plan.n_rows], row_a[: plan.n_rows]], 1), td.bra_rows)
    assert plan.qtotal == td.qtotal
    # block table == the loop of to_npc_array
    qc = 1 if mode == 0 else -1
    want = []
    for q in np.unique(td.q_ket):
        kr = np.flatnonzero(td.q_ket == q); br = np.flatnonzero(td.q_bra == q + td.qtotal * qc)
        if br.size:
            want.append([br[0], br.size, kr[0], kr.size, int(td.sets_ket[kr[0]].sum()), q])
    assert np.array_equal(blocks[: 6 * plan.n_blocks].reshape(-1, 6), np.array(want))
