"""Gutzwiller projection (reference gutzwiller.py:95-486): brute-force known answers on the dense state,
the device path on the CPU simulator and on the GPU."""
import itertools

import numpy as np
import pytest

import slater_oracle as so
from temfpy_b200 import gutzwiller as gw, slater
from tests import helpers


def fermion_state(H, spinful, tp):
    C_, _ = so.correlation_matrix(H)
    return C_, so.mps_to_state(so.C_to_MPS(C_, tp, spinful=spinful))


def brute_force(psi, kind):
    """amplitudes of the fermion state restricted to the allowed pair occupations (gutzwiller.py:104-117 /
    :294-303); spin index order as documented in temfpy_b200.gutzwiller."""
    L = psi.ndim // 2
    occ = {"simple": {0: (1, 0), 1: (0, 1)},      # abrikosov:    up, down
           "PH": {0: (0, 0), 1: (1, 1)}}[kind]     # abrikosov_ph: down, up
    out = np.zeros((2,) * L, dtype=complex)
    for s in itertools.product((0, 1), repeat=L):
        out[s] = psi[tuple(x for sj in s for x in occ[sj])]
    return out


def spin_ops(kind):
    up, dn = (0, 1) if kind == "simple" else (1, 0)
    Sz = np.zeros((2, 2)); Sz[up, up], Sz[dn, dn] = 0.5, -0.5
    Sp = np.zeros((2, 2)); Sp[up, dn] = 1.0
    return Sz, Sp, Sp.T


def heisenberg_bonds(phi, kind):
    L = phi.ndim
    phi = phi / np.linalg.norm(phi)
    Sz, Sp, Sm = spin_ops(kind)

    def apply(op, i, v):
        return np.moveaxis(np.tensordot(op, v, axes=(1, i)), 0, i)
    out = []
    for i in range(L - 1):
        e = np.vdot(phi, apply(Sz, i, apply(Sz, i + 1, phi)))
        e += 0.5 * np.vdot(phi, apply(Sp, i, apply(Sm, i + 1, phi)))
        e += 0.5 * np.vdot(phi, apply(Sm, i, apply(Sp, i + 1, phi)))
        out.append(e.real)
    return np.array(out)


def spin_state(m):
    """dense state of a finite spin BlockMPS (right-canonical or bare)."""
    psi = np.ones((1, 1), dtype=complex)
    for i in range(m.L):
        T = m.get_B_dense(i)
        psi = np.tensordot(psi, T, axes=(1, 0)).reshape(-1, T.shape[2])
    return psi.reshape((2,) * m.L)


def test_brute_force_known_answers():
    """both conventions give the same singlet with the listed bond energies."""
    H = so.hopping_chain(6)
    tp = {"chi_max": 4096, "svd_min": 1e-7}
    want = np.array([-0.7024, -0.2055, -0.6640, -0.2055, -0.7024])
    for kind in ("simple", "PH"):
        _, psi = fermion_state(H, kind, tp)
        phi = brute_force(psi, kind)
        assert abs(np.linalg.norm(phi) ** 2 - 0.084184) < 1e-6
        assert np.abs(heisenberg_bonds(phi, kind) - want).max() < 1e-4


def _check(be, L, kind, tp, canonical):
    H = so.hopping_chain(L)
    C_, psi = fermion_state(H, kind, tp)
    phi = brute_force(psi, kind)
    fm = slater.C_to_MPS(C_, tp, spinful=kind, _backend=be, as_tenpy=False)
    fn = gw.abrikosov if kind == "simple" else gw.abrikosov_ph
    sm = fn(fm, return_canonical=canonical, _backend=be)
    got = spin_state(sm)
    ov = abs(np.vdot(phi, got)) / (np.linalg.norm(phi) * np.linalg.norm(got))
    assert ov > 1 - 1e-10, ov
    if canonical:
        assert abs(np.linalg.norm(got) - 1) < 1e-12
        for i in range(sm.L):       # right-canonical
            T = sm.get_B_dense(i)
            e = np.einsum("apc,bpc->ab", T, T.conj())
            assert np.abs(e - np.eye(len(e))).max() < 1e-12
        for lam in sm.lams:
            assert abs(np.linalg.norm(lam) - 1) < 1e-12
        if kind == "PH":            # 2Sz charges: q(vL) + q_p = q(vR), total Sz = 0
            qp = np.array([-1, 1])
            for i in range(sm.L):
                T = sm.get_B_dense(i)
                a, p, b = np.nonzero(np.abs(T) > 1e-14)
                assert np.all(sm.charges[i][a] + qp[p] == sm.charges[i + 1][b])
            assert np.all(sm.charges[0] == 0) and np.all(sm.charges[sm.L] == 0)
    else:
        assert abs(np.linalg.norm(got) - np.linalg.norm(phi)) < 1e-10 * np.linalg.norm(phi) + 1e-13
    return sm


@pytest.mark.parametrize("L,kind,canonical", [(6, "simple", True), (6, "PH", True), (8, "PH", False),
                                              (8, "simple", False), (10, "PH", True)])
def test_sim_projection_vs_brute_force(sim_backend, L, kind, canonical):
    if L % 2 and kind == "simple":
        pytest.skip("odd L has no half filling")
    _check(sim_backend, L, kind, {"chi_max": 4096, "svd_min": 1e-7}, canonical)


def test_sim_ortho_center_inside_pair(sim_backend):
    """Schmidt values of the centre bond inside a pair (odd ortho_center)."""
    H = so.hopping_chain(6)
    tp = {"chi_max": 4096, "svd_min": 1e-7}
    C_, psi = fermion_state(H, "PH", tp)
    phi = brute_force(psi, "PH")
    for oc in (5, 4, 7):
        fm = slater.C_to_MPS(C_, tp, spinful="PH", ortho_center=oc, _backend=sim_backend, as_tenpy=False)
        got = spin_state(gw.abrikosov_ph(fm, _backend=sim_backend))
        assert abs(np.vdot(phi, got)) / np.linalg.norm(phi) > 1 - 1e-10


def _device_vs_host_canonical(be, L, kind, tp, monkeypatch):
    """The device sweep (tmf_canon_*) against the host sweep: same Schmidt spectrum on every bond (1e-12), same bond
    dimensions and charges, same state."""
    C_, _ = so.correlation_matrix(so.hopping_chain(L))
    fn = gw.abrikosov if kind == "simple" else gw.abrikosov_ph
    fm = slater.C_to_MPS(C_, tp, spinful=kind, _backend=be, as_tenpy=False)
    monkeypatch.setattr(gw, "CANONICAL_FORM", "device")
    dev = fn(fm, _backend=be)
    if dev.L <= 10:                      # (the device result against the brute-force projection of the dense state)
        _, psi = fermion_state(so.hopping_chain(L), kind, tp)
        phi, got = brute_force(psi, kind), spin_state(dev)
        assert abs(np.vdot(phi, got)) / np.linalg.norm(phi) > 1 - 1e-10
    monkeypatch.setattr(gw, "CANONICAL_FORM", "host")
    host = fn(fm, _backend=be)
    assert dev.meta["canonical_form"] == "device" and host.meta["canonical_form"] == "host"
    for x in range(dev.L + 1):
        a, b = np.sort(dev.lams[x])[::-1], np.sort(host.lams[x])[::-1]
        assert len(a) == len(b), (x, len(a), len(b))
        assert np.abs(a - b).max() < 1e-12, (x, np.abs(a - b).max())
        assert np.array_equal(np.sort(dev.charges[x]), np.sort(host.charges[x]))
    E = np.ones((1, 1))
    for i in range(dev.L):
        E = np.einsum("ab,apc,bpd->cd", E, dev.get_B_dense(i), host.get_B_dense(i), optimize=True)
    assert abs(abs(E[0, 0]) - 1) < 1e-10, E
    return dev


@pytest.mark.parametrize("L,kind,chi", [(8, "PH", 4096), (10, "simple", 4096), (12, "PH", 24), (12, "simple", 24),
                                        (20, "PH", 24)])
def test_sim_device_canonical_form(sim_backend, L, kind, chi, monkeypatch):
    _device_vs_host_canonical(sim_backend, L, kind, {"chi_max": chi, "svd_min": 1e-7}, monkeypatch)


def _resident_vs_staged(be, L, kind, tp):
    """The projection reads the fermion blocks where the conversion left them (no staging) and gives the same
    numbers, bit for bit, as the path that stages host-resident blocks."""
    C_, _ = so.correlation_matrix(so.hopping_chain(L))
    fn = gw.abrikosov if kind == "simple" else gw.abrikosov_ph
    fm = slater.C_to_MPS(C_, tp, spinful=kind, _backend=be, as_tenpy=False)
    a = fn(fm, return_canonical=False, _backend=be)
    assert a.meta["resident_operands"] == 2 * a.meta["gemm_jobs"] > 0 and a.meta["staged_elems"] == 0
    fm2 = slater.C_to_MPS(C_, tp, spinful=kind, _backend=be, as_tenpy=False, _keep_device=False)
    b = fn(fm2, return_canonical=False, _backend=be)
    assert b.meta["resident_operands"] == 0 and b.meta["staged_elems"] > 0
    for i in range(a.L):
        assert np.array_equal(a.get_B_dense(i), b.get_B_dense(i))
    return a


@pytest.mark.parametrize("kind", ["simple", "PH"])
def test_sim_resident_blocks(sim_backend, kind):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _resident_vs_staged(sim_backend, 8, kind, {"chi_max": 4096, "svd_min": 1e-7})


def _pf_projection(be, kind, canonical):
    """Parity-conserving input (pfaffian.C_to_MPS, complex tensors): the projection of the MPS equals the
    brute-force projection of its own dense state (gutzwiller.py:171-172 / :364-367)."""
    import pfaffian_oracle as po
    from temfpy_b200 import pfaffian as pf
    fn = gw.abrikosov if kind == "simple" else gw.abrikosov_ph
    done = 0
    for seed in range(12):
        Cm = po.correlation_matrix(po.random_bdg(6, 100 + seed), "C->C")
        fm = pf.C_to_MPS(Cm, {"chi_max": 4096, "svd_min": 1e-7}, basis="C", _backend=be, as_tenpy=False)
        total = int(np.asarray(fm.charges[fm.L]).ravel()[0])
        if total % 2 != ((fm.L // 2) % 2 if kind == "simple" else 0):
            with pytest.raises(AssertionError):
                fn(fm, _backend=be)
            continue
        psi = so.mps_to_state(helpers.block_mps_to_dense(fm))
        phi = brute_force(psi, kind)
        sm = fn(fm, return_canonical=canonical, _backend=be)
        assert sm.conserve is None and all(np.all(q == 0) for q in sm.charges)
        got = spin_state(sm)
        ov = abs(np.vdot(phi, got)) / (np.linalg.norm(phi) * np.linalg.norm(got))
        assert ov > 1 - 1e-10, (seed, ov)
        if not canonical:
            assert abs(np.linalg.norm(got) - np.linalg.norm(phi)) < 1e-10 * np.linalg.norm(phi) + 1e-13
        done += 1
        if done == 2:
            break
    assert done > 0


@pytest.mark.parametrize("kind,canonical", [("simple", False), ("PH", False), ("PH", True)])
def test_sim_parity_conserving_input(sim_backend, kind, canonical):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _pf_projection(sim_backend, kind, canonical)


def test_sim_validation(sim_backend):
    C_, _ = so.correlation_matrix(so.hopping_chain(5))
    fm = slater.C_to_MPS(C_, {"chi_max": 64}, _backend=sim_backend, as_tenpy=False)
    with pytest.raises(AssertionError):
        gw.abrikosov(fm, _backend=sim_backend)            # odd length
    C_, n = so.correlation_matrix(so.hopping_chain(6), N=2)
    fm = slater.C_to_MPS(C_, {"chi_max": 64}, _backend=sim_backend, as_tenpy=False)
    with pytest.raises(AssertionError):
        gw.abrikosov(fm, _backend=sim_backend)            # 2 fermions on 3 spin sites
    # inplace=True rewrites the given BlockMPS and returns nothing (gutzwiller.py:210 / :280)
    C_, _ = so.correlation_matrix(so.hopping_chain(6))
    fm = slater.C_to_MPS(C_, {"chi_max": 64}, spinful="PH", _backend=sim_backend, as_tenpy=False)
    want = gw.abrikosov_ph(fm, _backend=sim_backend)
    assert gw.abrikosov_ph(fm, inplace=True, _backend=sim_backend) is None
    assert fm.site_type == "SpinHalfSite" and fm.L == want.L and fm.conserve == "Sz"
    assert all(np.array_equal(fm.get_B_dense(i), want.get_B_dense(i)) for i in range(fm.L))


@pytest.mark.gpu
@pytest.mark.parametrize("L,kind", [(8, "simple"), (8, "PH"), (10, "PH")])
def test_gpu_projection_vs_brute_force(gpu_backend, L, kind):
    _check(gpu_backend, L, kind, {"chi_max": 4096, "svd_min": 1e-7}, True)


@pytest.mark.gpu
@pytest.mark.parametrize("L", [64, 256])
def test_gpu_cfg3_heisenberg(gpu_backend, L):
    """BASELINE configs[2] (L = 256 spin sites from 512 fermion sites, chi = 256) and a shorter chain: the
    Gutzwiller-projected half-filled Fermi sea; both conventions must give the same spin state,
    a reference-independent check at the size where the dense brute force is impossible."""
    tp = {"chi_max": 256}
    H = so.hopping_chain(L)
    C_, _ = so.correlation_matrix(H)
    a = gw.abrikosov(slater.C_to_MPS(C_, tp, spinful="simple", _backend=gpu_backend, as_tenpy=False), _backend=gpu_backend)
    b = gw.abrikosov_ph(slater.C_to_MPS(C_, tp, spinful="PH", _backend=gpu_backend, as_tenpy=False), _backend=gpu_backend)
    # overlap of the two spin MPS; index maps: abrikosov 0 = up, abrikosov_ph 1 = up
    E = np.ones((1, 1))
    for i in range(L):
        E = np.einsum("ab,apc,bpd->cd", E, a.get_B_dense(i), b.get_B_dense(i)[:, ::-1, :], optimize=True)
    # (L = 256: chi = 256 truncates the critical 512-site fermion chain; the two conversions discard different
    #  states, measured overlap 1 - 2.2e-6)
    assert abs(abs(E[0, 0]) - 1) < (1e-6 if L <= 64 else 1e-5), E
    assert max(a.chi) <= 256 and max(b.chi) <= 256
    print("cfg3-like chi_proj:", max(a.chi), max(b.chi), "overlap", abs(E[0, 0]))


@pytest.mark.gpu
@pytest.mark.parametrize("L,kind,chi", [(32, "PH", 64), (32, "simple", 64), (96, "PH", 128)])
def test_gpu_device_canonical_form(gpu_backend, L, kind, chi, monkeypatch):
    _device_vs_host_canonical(gpu_backend, L, kind, {"chi_max": chi}, monkeypatch)


@pytest.mark.gpu
def test_gpu_resident_blocks_and_parity_input(gpu_backend):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for kind in ("simple", "PH"):
            _resident_vs_staged(gpu_backend, 48, kind, {"chi_max": 128})
        _pf_projection(gpu_backend, "PH", True)
        _pf_projection(gpu_backend, "simple", False)


def _project_kernel_vs_numpy(be, cplx, seed):
    """tmf_gutzwiller_project through the C ABI on random jobs: both storage orders of either operand (element
    strides), the three optional scalings, f64 and c128, output blocks scattered inside a larger tensor -- against
    NumPy, and the zero fill of everything the jobs do not touch."""
    import ctypes as C
    from temfpy_b200 import _lib
    rng = np.random.default_rng(seed)
    es = 2 if cplx else 1
    dt = np.complex128 if cplx else np.float64

    def rnd(*shape):
        x = rng.normal(size=shape)
        return x + 1j * rng.normal(size=shape) if cplx else x
    pool, off, specs = [], 0, []
    n_out_rows, n_out_cols = 300, 2 * 170            # one "spin tensor": rows vL, columns (s, vR) -> so_i = 2 * 170
    want = np.zeros((n_out_rows, n_out_cols), dtype=dt)
    r0 = 0
    for u, (m, k, n) in enumerate([(1, 1, 1), (7, 3, 5), (64, 16, 64), (65, 17, 66), (130, 40, 3), (20, 70, 90)]):
        ta, tb = bool(rng.integers(2)), bool(rng.integers(2))
        A, B = rnd(m, k), rnd(k, n)
        As, Bs = (A.T.copy() if ta else A.copy()), (B.T.copy() if tb else B.copy())
        ks = rng.uniform(0.5, 1.5, size=k) if u % 2 == 0 else None
        rs = rng.uniform(0.5, 1.5, size=m) if u % 3 == 0 else None
        cs = rng.uniform(0.5, 1.5, size=n) if u % 3 == 1 else None
        s_idx, c0 = u % 2, int(rng.integers(0, 170 - n + 1))
        ref = (A * (ks[None, :] if ks is not None else 1.0)) @ B
        if rs is not None:
            ref = ref * rs[:, None]
        if cs is not None:
            ref = ref * cs[None, :]
        want[r0: r0 + m, s_idx * 170 + c0: s_idx * 170 + c0 + n] = ref
        specs.append(dict(m=m, k=k, n=n, ta=ta, tb=tb, a=off, b=off + As.size, ks=ks, rs=rs, cs=cs,
                          out=r0 * n_out_cols + s_idx * 170 + c0))
        pool += [As.ravel(), Bs.ravel()]
        off += As.size + Bs.size
        r0 += m
    buf = be.from_host(np.concatenate(pool).astype(dt).view(np.float64))
    scal = [v for sp in specs for v in (sp["ks"], sp["rs"], sp["cs"]) if v is not None]
    sc_d = be.from_host(np.concatenate(scal))
    outd = be.from_host(np.full(es * n_out_rows * n_out_cols, 7.0))          # must be overwritten by the zero fill
    jobs = (_lib.GutzJob * len(specs))()
    so_ = 0
    for g, sp in zip(jobs, specs):
        g.A, g.B = be.ptr(buf) + 8 * es * sp["a"], be.ptr(buf) + 8 * es * sp["b"]
        g.sa_i, g.sa_k = (1, sp["m"]) if sp["ta"] else (sp["k"], 1)
        g.sb_k, g.sb_n = (1, sp["k"]) if sp["tb"] else (sp["n"], 1)
        g.m, g.k, g.n = sp["m"], sp["k"], sp["n"]
        g.out, g.so_i = be.ptr(outd) + 8 * es * sp["out"], n_out_cols
        for name in ("ks", "rs", "cs"):
            if sp[name] is not None:
                setattr(g, {"ks": "k_scale", "rs": "row_scale", "cs": "col_scale"}[name], be.ptr(sc_d) + 8 * so_)
                so_ += len(sp[name])
    desc = be.empty(int(be.lib.tmf_gutz_desc_bytes(jobs, len(specs))), np.uint8)
    _lib.check(be.lib, be.lib.tmf_gutzwiller_project(jobs, len(specs), int(cplx), be.ptr(outd),
                                                     8 * es * n_out_rows * n_out_cols, be.ptr(desc), be.stream))
    be.sync()
    got = be.to_host(outd, es * n_out_rows * n_out_cols)
    got = (got.view(np.complex128) if cplx else got).reshape(n_out_rows, n_out_cols)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    assert np.array_equal(got == 0, want == 0)


@pytest.mark.parametrize("cplx", [False, True])
def test_sim_project_kernel_vs_numpy(sim_backend, cplx):
    _project_kernel_vs_numpy(sim_backend, cplx, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("cplx,seed", [(False, 1), (True, 2), (False, 3)])
def test_gpu_project_kernel_vs_numpy(gpu_backend, cplx, seed):
    _project_kernel_vs_numpy(gpu_backend, cplx, seed)


# ---------------------------------------------------------------------------------------------
# infinite MPS input (gutzwiller.py:197-205, :236-244, :268 / :473)
# ---------------------------------------------------------------------------------------------
def _heis_two_site(Ba, Bb, lam_left, kind):
    """<S_i . S_{i+1}> of two neighbouring right-canonical tensors with the Schmidt values on their left bond."""
    Sz, Sp, Sm = spin_ops(kind)
    theta = np.einsum("a,apb,bqc->apqc", lam_left, Ba, Bb)
    val = 0.0
    for Oa, Ob, w in ((Sz, Sz, 1.0), (Sp, Sm, 0.5), (Sm, Sp, 0.5)):
        val += w * np.einsum("apqc,rp,sq,arsc->", theta.conj(), Oa, Ob, theta).real
    return val / np.einsum("apqc,apqc->", theta.conj(), theta).real


def _infinite_vs_finite(be, kind, q_left=0):
    """The Gutzwiller-projected unit cell of slater.C_to_iMPS against the centre of a long finite chain projected the
    same way (gapped, dimerised chain: finite-size and truncation effects are exponentially small): canonical-form
    conditions, and the nearest-neighbour <S.S> on the two inequivalent bonds."""
    from tests.test_imps import dimer_chain
    fn = gw.abrikosov if kind == "simple" else gw.abrikosov_ph
    kw = dict(q_left=q_left) if kind == "simple" else {}
    tp = {"chi_max": 64, "svd_min": 1e-6}
    Ls, cell, cut = 28, 2, 14
    Cs, _ = so.correlation_matrix(dimer_chain(Ls, tnn=0.0))
    Cl, _ = so.correlation_matrix(dimer_chain(Ls + cell, tnn=0.0))
    im, _err = slater.C_to_iMPS(Cs, Cl, tp, cell, cut, spinful=kind, _backend=be, as_tenpy=False)
    bare = fn(im, return_canonical=False, _backend=be, **kw)
    sp = fn(im, return_canonical=True, _backend=be, **kw)
    assert sp.bc == "infinite" and sp.L == cell and sp.form == ["B"] * cell
    B = [sp.get_B_dense(i) for i in range(sp.L)]
    for i in range(sp.L):           # right-canonical, and the left environment lambda^2 is carried to the next bond
        e = np.einsum("apc,bpc->ab", B[i], B[i].conj())
        assert np.abs(e - np.eye(len(e))).max() < 1e-9, (i, np.abs(e - np.eye(len(e))).max())
        le = np.einsum("apb,a,apc->bc", B[i].conj(), sp.lams[i] ** 2, B[i])
        assert np.abs(le - np.diag(sp.lams[i + 1] ** 2)).max() < 1e-9, i
        assert abs(np.linalg.norm(sp.lams[i]) - 1) < 1e-12
    # same state as the bare projected cell: mixed transfer matrix
    from tests.test_imps import cell_transfer_eig
    A = [bare.get_B_dense(i) for i in range(bare.L)]
    fid = abs(cell_transfer_eig(A, B)) ** 2 / abs(cell_transfer_eig(A, A) * cell_transfer_eig(B, B))
    assert fid > 1 - 1e-9, fid
    if kind == "PH":                # 2 Sz charges: q(vL) + q_p - qtotal = q(vR), bond L = bond 0
        qp = np.array([-1, 1])
        for i in range(sp.L):
            a, p, b = np.nonzero(np.abs(B[i]) > 1e-13)
            assert np.all(sp.charges[i][a] + qp[p] - sp.tensors[i].qtotal == sp.charges[i + 1][b])
        assert np.array_equal(sp.charges[0], sp.charges[sp.L])
    # against the centre of a finite chain
    Lf = 40
    Cf, _ = so.correlation_matrix(dimer_chain(Lf, tnn=0.0))
    fin = fn(slater.C_to_MPS(Cf, tp, spinful=kind, _backend=be, as_tenpy=False), _backend=be)
    Bf = [fin.get_B_dense(i) for i in range(fin.L)]
    j = Lf // 2                     # even site: same position inside the dimer as cell site 0
    want = [_heis_two_site(Bf[j], Bf[j + 1], fin.lams[j], kind), _heis_two_site(Bf[j + 1], Bf[j + 2], fin.lams[j + 1], kind)]
    got = [_heis_two_site(B[0], B[1], sp.lams[0], kind), _heis_two_site(B[1], B[0], sp.lams[1], kind)]
    assert np.abs(np.array(got) - np.array(want)).max() < 1e-4, (got, want)
    return got, want


@pytest.mark.parametrize("kind", ["simple", "PH"])
def test_sim_infinite_input(sim_backend, kind):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _infinite_vs_finite(sim_backend, kind)


def test_sim_infinite_validation(sim_backend):
    from tests.test_imps import dimer_chain
    import warnings
    Cs, _ = so.correlation_matrix(dimer_chain(20, tnn=0.0))
    Cl, _ = so.correlation_matrix(dimer_chain(22, tnn=0.0))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        im, _ = slater.C_to_iMPS(Cs, Cl, {"chi_max": 32}, 2, 10, spinful="simple", _backend=sim_backend, as_tenpy=False)
    with pytest.raises(ValueError):
        gw.abrikosov(im, _backend=sim_backend)                      # q_left is mandatory (gutzwiller.py:199-200)
    with pytest.raises(ValueError):
        gw.abrikosov(im, q_left=99, _backend=sim_backend)           # not a sector of the leftmost leg (:201-205)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["simple", "PH"])
def test_gpu_infinite_input(gpu_backend, kind):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        print(_infinite_vs_finite(gpu_backend, kind))


def test_sim_infinite_parity_input(sim_backend):
    """Unit cell of pfaffian.C_to_iMPS (parity-conserving, complex): the projected cell tensor equals the dense
    contraction of the two fermion tensors under the three masks (gutzwiller.py:236-244 with parity masks)."""
    import warnings
    import pfaffian_oracle as po
    from temfpy_b200 import pfaffian as pf
    Ls, cell, cut = 20, 2, 10
    Cs = po.correlation_matrix(po.bdg_chain(Ls, mu=0.3, delta=0.4), "C->C")
    Cl = po.correlation_matrix(po.bdg_chain(Ls + cell, mu=0.3, delta=0.4), "C->C")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        im, _ = pf.C_to_iMPS(Cs, Cl, {"chi_max": 32}, cell, cut, basis="C", _backend=sim_backend, as_tenpy=False)
        th = np.einsum("apb,bqc->apqc", im.get_B_dense(0), im.get_B_dense(1))
        cL, cR = np.asarray(im.charges[0]), np.asarray(im.charges[2])
        for q_left in (0, 1):
            sp = gw.abrikosov(im, q_left=q_left, return_canonical=False, _backend=sim_backend)
            ref = np.stack([th[:, 1, 0, :], th[:, 0, 1, :]], axis=1)[cL % 2 == q_left][:, :, cR % 2 == q_left]
            assert sp.bc == "infinite" and sp.L == 1 and np.array_equal(sp.get_B_dense(0), ref)
        with pytest.raises(AssertionError):
            gw.abrikosov_ph(im, _backend=sim_backend)          # odd parity per cell (gutzwiller.py:370-372)


def _canon_abi_vs_host(be, seed):
    """tmf_canon_* through the C ABI on a random charge-conserving MPS with rank-deficient, tall and wide blocks
    (bond dimensions that grow and shrink, a dead sector): same Schmidt spectra as the host sweep, right-canonical
    tensors, same state."""
    import ctypes as C
    from temfpy_b200 import _lib
    rng = np.random.default_rng(seed)
    qp = np.array([-1, 1])
    dims = [{0: 1}, {-1: 1, 1: 1}, {-2: 3, 0: 5, 2: 2}, {-3: 2, -1: 7, 1: 9, 3: 1}, {-2: 12, 0: 4, 2: 6},
            {-1: 3, 1: 3}, {0: 2}, {-1: 1, 1: 1}, {0: 1}]
    qs = [np.concatenate([np.full(n, q) for q, n in sorted(d.items())]).astype(np.int64) for d in dims]
    L = len(dims) - 1
    T = []
    for j in range(L):
        a, b = len(qs[j]), len(qs[j + 1])
        t = np.zeros((a, 2, b))
        for s in range(2):
            mask = (qs[j][:, None] + qp[s]) == qs[j + 1][None, :]
            t[:, s, :] = rng.normal(size=(a, b)) * mask
        if j == 3:                       # a rank-deficient block: two equal columns inside one sector
            t[:, :, 1] = t[:, :, 0]
        T.append(t)
    ref_T, ref_l, ref_q = gw._canonical_form_finite(T, qs, qp, 1e-12)
    flat = np.concatenate([t.ravel() for t in T])
    offs = np.concatenate(([0], np.cumsum([t.size for t in T])))[:-1]
    got = gw._canonical_form_device(be, (be.from_host(flat), [int(o) for o in offs]), qs, qp, 1e-12)
    assert got is not None
    dT, dl, dq = got
    for x in range(L + 1):
        a, b = np.sort(dl[x])[::-1], np.sort(ref_l[x])[::-1]
        assert len(a) == len(b) and np.abs(a - b).max() < 1e-12, (x, a, b)
        assert np.array_equal(np.sort(dq[x]), np.sort(ref_q[x]))
    for t in dT:
        e = np.einsum("apc,bpc->ab", t, t)
        assert np.abs(e - np.eye(len(e))).max() < 1e-12
    E = np.ones((1, 1))
    for a, b in zip(dT, ref_T):
        E = np.einsum("ab,apc,bpd->cd", E, a, b, optimize=True)
    assert abs(abs(E[0, 0]) - 1) < 1e-12


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_sim_canon_abi_vs_host(sim_backend, seed):
    _canon_abi_vs_host(sim_backend, seed)


@pytest.mark.gpu
def test_gpu_canon_abi_vs_host(gpu_backend):
    for seed in (3, 4):
        _canon_abi_vs_host(gpu_backend, seed)
