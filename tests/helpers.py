"""Shared helpers of the test-suite (comparison rules are documented in DESIGN.md "Parity")."""
import os

import numpy as np

import slater_oracle as so
from temfpy_b200 import engine
from temfpy_b200.schmidt_utils import to_stopping_condition

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def random_hamiltonian(L, seed, decay=2.0):
    """Pattern of the reference's examples/slater.py:15-20."""
    rng = np.random.default_rng(seed)
    H = rng.normal(size=(L, L))
    H = H + H.T
    d = np.abs(np.subtract.outer(np.arange(L), np.arange(L)))
    return H * np.exp(-d / decay)


def cylinder_hamiltonian(Lx, Ly, t=-1.0):
    """cfg4: square lattice, site = x*Ly + y, periodic in y, open in x."""
    L = Lx * Ly
    H = np.zeros((L, L))
    for x in range(Lx):
        for y in range(Ly):
            i = x * Ly + y
            j = x * Ly + (y + 1) % Ly
            H[i, j] = H[j, i] = t
            if x + 1 < Lx:
                j = (x + 1) * Ly + y
                H[i, j] = H[j, i] = t
    return H


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_trunc(g):
    chi = int(g["chi_max"])
    return {"chi_max": None if chi < 0 else chi, "svd_min": float(g["svd_min"])}


def golden_dense_mps(g) -> so.DenseMPS:
    """Dense MPS assembled from the reference's own block values stored in a fixture."""
    L, oc = int(g["L"]), int(g["oc"])
    tensors, lams, charges = [], [], []
    for x in range(L + 1):
        lam = g[f"bond{x}_lam"]
        lams.append(lam / np.linalg.norm(lam))
        charges.append(g[f"bond{x}_charge"])
    for i in range(L):
        mode = "right" if i >= oc else "left"
        chi_bra = len(lams[i + 1]) if mode == "right" else len(lams[i])
        chi_ket = len(lams[i]) if mode == "right" else len(lams[i + 1])
        q_alpha = charges[i + 1] if mode == "right" else charges[i]
        # bra rows in the reference order: [p=0 | p=1], stably sorted by pipe charge
        p = np.repeat([0, 1], chi_bra)
        a = np.tile(np.arange(chi_bra), 2)
        q = q_alpha[a] + (p if mode == "left" else -p)
        order = np.argsort(q, kind="stable")
        p, a = p[order], a[order]
        T = np.zeros((2, chi_bra, chi_ket), dtype=g[f"site{i}_S"].dtype)
        for b in range(int(g[f"site{i}_nblocks"])):
            _, r0, nr, c0, nc = (int(v) for v in g[f"site{i}_blk{b}_meta"])
            rows = slice(r0, r0 + nr)
            T[p[rows][:, None], a[rows][:, None], np.arange(c0, c0 + nc)[None, :]] = g[f"site{i}_blk{b}"]
        tensors.append(np.transpose(T, (2, 0, 1)) if mode == "right" else np.transpose(T, (1, 0, 2)))
    return so.DenseMPS(tensors=tensors, lams=lams, charges=charges,
                       form=["A"] * oc + ["B"] * (L - oc), ortho_center=oc)


def chain_to_dense(res: engine.ChainResult) -> so.DenseMPS:
    L = res.L
    lams = [res.bonds[x].schmidt_values / np.linalg.norm(res.bonds[x].schmidt_values) for x in range(L + 1)]
    return so.DenseMPS(tensors=[res.sites[i].dense() for i in range(L)], lams=lams,
                       charges=[res.bonds[x].charge for x in range(L + 1)],
                       form=["A"] * res.ortho_center + ["B"] * (L - res.ortho_center),
                       ortho_center=res.ortho_center)


def run_native(backend, C, trunc, N=None, **kw) -> engine.ChainResult:
    L = len(C)
    if N is None:
        N = int(round(np.trace(C)))
    Cd = backend.from_host(np.ascontiguousarray(C, dtype=np.float64).ravel())
    return engine.run_chain(backend, Cd, L, L, to_stopping_condition(trunc), N, **kw)


def default_policy(C):
    """The options the public entry points choose for this input (slater.C_to_MPS)."""
    return dict(snap=engine.snap_policy(C))


def ambiguous_bonds(ref: so.DenseMPS, trunc, margin=1e-6):
    """Threshold-margin audit: bonds whose truncation decision lies inside the
    eigenvalue noise of the reference itself (a near-degenerate multiplet straddles the chi_max /
    svd_min cut).  There the reference's kept set is decided by LAPACK rounding noise and is not
    reproducible by *any* other solver; such bonds are compared through their Schmidt spectrum only.
    """
    tp = so.Trunc.make(trunc)
    out = set()
    for x, lam in enumerate(ref.lams):
        if len(lam) < 2:
            continue
        lv = np.sort(-np.log(lam / lam.max()))
        at_chi = tp.chi_max is not None and len(lam) >= tp.chi_max - 8
        at_svd = lv[-1] > -np.log(tp.svd_min) - 1e-3
        tail_deg = np.any(np.diff(lv[-6:]) < margin)
        if (at_chi and tail_deg) or at_svd:
            out.add(x)
    return out


def compare_mps(ref: so.DenseMPS, got: so.DenseMPS, trunc, lam_abs=1e-8, noise=None, ent_tol=1e-10, ov_tol=1e-10,
                check_overlap=True):
    """The parity gate of BASELINE.json: identical bond dimensions and charge sectors (integers,
    exact, outside the audited ambiguous bonds), Schmidt values, entropies, overlap."""
    amb = ambiguous_bonds(ref, trunc)
    if noise is None:       # eigenvalue rounding noise of an n x n symmetric eigenproblem, n ~ L/2
        noise = 4e-15 * np.sqrt(ref.L)
    report = dict(ambiguous=sorted(amb), lam_rel=0.0, lam_abs=0.0)
    for x in range(ref.L + 1):
        a, b = ref.lams[x], got.lams[x]
        if x in amb:
            # everything clearly above the contested multiplet must agree as a multiset
            # (relative to the largest value: the normalisation depends on which multiplet was kept)
            a, b = a / a.max(), b / b.max()
            cutv = max(a.min(), b.min()) * (1 + 1e-3)
            sa, sb = np.sort(a[a > cutv])[::-1], np.sort(b[b > cutv])[::-1]
            assert len(sa) == len(sb) and np.all(np.abs(sa - sb) <= 1e-12 * sa + np.minimum(noise / (2 * sa), lam_abs)), \
                f"bond {x}: spectrum differs"
            continue
        assert len(a) == len(b), f"bond {x}: chi {len(b)} != reference {len(a)}"
        assert np.array_equal(ref.charges[x], got.charges[x]), f"bond {x}: charge sectors differ"
        report["lam_abs"] = max(report["lam_abs"], float(np.max(np.abs(a - b))))
        # Tolerance model: lambda^2 is a product of mode eigenvalues e_i that *both*
        # solvers (LAPACK in the reference, Jacobi/Rayleigh-Ritz here) deliver with an absolute error
        # of a few 1e-15 * ||C||.  Hence |d(lambda^2)| <= noise, i.e. |d lambda| <= noise / (2 lambda):
        # 1e-12 relative holds for the well-conditioned values, the weak ones carry the reference's
        # own rounding noise.  `noise` ~ 4e-15 sqrt(L) was calibrated against 40-digit arithmetic (L = 20).
        tol = 1e-12 * a + np.minimum(noise / (2 * a), lam_abs)
        assert np.all(np.abs(a - b) <= tol), f"bond {x}: Schmidt values differ by {np.max(np.abs(a - b) / tol)} tol"
        big = a > 0.05 * a.max()
        if big.any():
            report["lam_rel"] = max(report["lam_rel"], float(np.max(np.abs(a[big] - b[big]) / a[big])))
    assert report["lam_rel"] < 1e-12, report
    ent = np.abs(so.entropies(ref.lams) - so.entropies(got.lams))
    keep = [x for x in range(ref.L + 1) if x not in amb]
    report["entropy"] = float(ent[keep].max()) if keep else 0.0
    assert report["entropy"] < ent_tol, report
    if check_overlap and not amb:
        o = abs(so.mps_overlap(ref, got)) / np.sqrt(abs(so.mps_overlap(ref, ref) * so.mps_overlap(got, got)))
        report["overlap"] = float(o)
        assert o >= 1 - ov_tol, report
    return report


# ---------------------------------------------------------------------------------------------
# Pfaffian path
# ---------------------------------------------------------------------------------------------
def block_mps_to_dense(m) -> so.DenseMPS:
    """temfpy_b200.mps.BlockMPS -> the oracle's dense container."""
    return so.DenseMPS(tensors=[m.get_B_dense(i) for i in range(m.L)], lams=list(m.lams), charges=list(m.charges),
                       form=list(m.form), ortho_center=m.ortho_center)


def golden_pf_mps(g) -> so.DenseMPS:
    """Dense MPS assembled from the reference's own tensors stored in a Pfaffian fixture."""
    L, oc = int(g["L"]), int(g["oc"])
    lams = [g[f"bond{x}_lam"] / np.linalg.norm(g[f"bond{x}_lam"]) for x in range(L + 1)]
    charges = [g[f"bond{x}_charge"] for x in range(L + 1)]
    tensors = []
    for i in range(L):
        T = g[f"site{i}_T"]
        tensors.append(np.transpose(T, (2, 0, 1)) if i >= oc else np.transpose(T, (1, 0, 2)))
    return so.DenseMPS(tensors=tensors, lams=lams, charges=charges, form=["A"] * oc + ["B"] * (L - oc),
                       ortho_center=oc)


def compare_pf_mps(ref: so.DenseMPS, got: so.DenseMPS, half_bonds=(), ent_tol=1e-10, ov_tol=1e-10):
    """Parity gate for the Pfaffian path: bond dimensions and parity sectors identical, Schmidt values within
    the tolerance model of compare_mps, entropies, normalised overlap.  On bonds that carry modes with
    eigenvalue exactly 1/2 the vacuum parity is a gauge choice (the reference fixes it through LAPACK's
    arbitrary basis of the degenerate eigenspace and a random shuffle, pfaffian.py:807-816, :867-874), so
    the charge table is compared as a multiset there."""
    L = ref.L
    noise = 4e-15 * np.sqrt(4 * L)
    rep = dict(lam_rel=0.0, lam_abs=0.0, gauge_bonds=0)
    for x in range(L + 1):
        a, b = ref.lams[x], got.lams[x]
        assert len(a) == len(b), f"bond {x}: chi {len(b)} != reference {len(a)}"
        if x in half_bonds:
            assert np.array_equal(np.sort(ref.charges[x]), np.sort(got.charges[x])), f"bond {x}: parity sectors differ"
            rep["gauge_bonds"] += int(not np.array_equal(ref.charges[x], got.charges[x]))
        else:
            assert np.array_equal(ref.charges[x], got.charges[x]), f"bond {x}: parity sectors differ"
        tol = 1e-12 * a + np.minimum(noise / (2 * a), 1e-8)
        assert np.all(np.abs(a - b) <= tol), f"bond {x}: Schmidt values differ by {np.max(np.abs(a - b) / tol)} tol"
        rep["lam_abs"] = max(rep["lam_abs"], float(np.max(np.abs(a - b))))
        big = a > 0.05 * a.max()
        rep["lam_rel"] = max(rep["lam_rel"], float(np.max(np.abs(a[big] - b[big]) / a[big])))
    assert rep["lam_rel"] < 1e-12, rep
    rep["entropy"] = float(np.abs(so.entropies(ref.lams) - so.entropies(got.lams)).max())
    assert rep["entropy"] < ent_tol, rep
    o = abs(so.mps_overlap(ref, got)) / np.sqrt(abs(so.mps_overlap(ref, ref) * so.mps_overlap(got, got)))
    rep["overlap"] = float(o)
    assert o >= 1 - ov_tol, rep
    return rep


# ---------------------------------------------------------------------------------------------
# full-size per-bond fixtures written by oracle/make_golden_full.py from the live reference
# ---------------------------------------------------------------------------------------------
def fixture_bond(g, x):
    """(k, filled_left, e, lam, masks, charge) of bond x as the reference computed them."""
    a, b = int(g["chi_off"][x]), int(g["chi_off"][x + 1])
    s0, s1 = int(g["sec_off"][x]), int(g["sec_off"][x + 1])
    e0, e1 = int(g["e_off"][x]), int(g["e_off"][x + 1])
    charge = np.repeat(g["sec_q"][s0:s1].astype(np.int64), g["sec_n"][s0:s1])
    return int(g["k"][x]), int(g["filled_left"][x]), g["e"][e0:e1], g["lam"][a:b], g["masks"][a:b], charge


def _lam_tol(a, noise, lam_abs=1e-8):
    """Tolerance model of compare_mps: 1e-12 relative plus the mode-eigenvalue rounding noise."""
    return 1e-12 * a + np.minimum(noise / (2 * a), lam_abs)


def compare_bonds_fixture(g, get_bond, noise=None, max_multiplet=16):
    """Integer + spectral parity of *every* bond against a reference-run fixture.

    A bond is `exact` when k, the filled count, chi, the sector table and every occupation mask are
    identical and the Schmidt values agree within the tolerance model.  Otherwise it must pass the
    noise audit: the reference's `truncate` (schmidt_utils.py:140-185) cut inside / next to a multiplet whose
    members are equal in exact arithmetic but differ by LAPACK rounding noise in the reference,
    i.e. (1) per charge sector the spectra above the contested multiplet agree as multisets, (2) everything
    below lies within 1 % of the cut, (3) at most `max_multiplet` values per side are contested.
    Returns the counts; raises AssertionError for a bond that is neither."""
    L = int(g["L"])
    if noise is None:
        noise = 4e-15 * np.sqrt(L)
    rep = dict(bonds=L + 1, exact=0, ambiguous=0, max_dchi=0, lam_rel=0.0, entropy=0.0, k_noise=0,
               ambiguous_list=[])
    for x in range(L + 1):
        k, fl, e, lam, masks, charge = fixture_bond(g, x)
        b = get_bond(x)
        bl = np.asarray(b.schmidt_values)
        bq = np.asarray(b.charge, dtype=np.int64)
        same = (len(bl) == len(lam) and b.k == k and b.filled_left == fl and np.array_equal(bq, charge)
                and np.array_equal(np.asarray(b.masks, dtype=np.uint64), masks))
        a_n, b_n = lam / np.linalg.norm(lam), bl / np.linalg.norm(bl)
        if same and np.all(np.abs(a_n - b_n) <= _lam_tol(a_n, noise)):
            rep["exact"] += 1
            big = a_n > 0.05 * a_n.max()
            rep["lam_rel"] = max(rep["lam_rel"], float(np.max(np.abs(a_n[big] - b_n[big]) / a_n[big])))
            rep["entropy"] = max(rep["entropy"], float(abs(so.entropies([a_n])[0] - so.entropies([b_n])[0])))
            continue
        # ---- noise audit ---------------------------------------------------------------------
        a_r, b_r = lam / lam.max(), bl / bl.max()
        cutv = max(a_r.min(), b_r.min()) * (1 + 1e-3)
        na, nb_ = int((a_r <= cutv).sum()), int((b_r <= cutv).sum())
        assert na <= max_multiplet and nb_ <= max_multiplet, f"bond {x}: {na}/{nb_} contested values"
        assert abs(len(lam) - len(bl)) <= max(na, nb_), f"bond {x}: chi {len(bl)} vs reference {len(lam)}"
        lowest = min(a_r.min(), b_r.min())
        assert lowest >= cutv * (1 - 1e-2), f"bond {x}: contested values spread below the cut ({lowest / cutv})"
        for q in np.union1d(charge, bq):
            sa = np.sort(a_r[(charge == q) & (a_r > cutv)])[::-1]
            sb = np.sort(b_r[(bq == q) & (b_r > cutv)])[::-1]
            assert len(sa) == len(sb), f"bond {x}, charge {q}: {len(sb)} vs {len(sa)} values above the cut"
            assert np.all(np.abs(sa - sb) <= _lam_tol(sa, noise) * 4), f"bond {x}, charge {q}: spectrum differs"
        rep["ambiguous"] += 1
        rep["ambiguous_list"].append(x)
        rep["max_dchi"] = max(rep["max_dchi"], abs(len(lam) - len(bl)))
        rep["k_noise"] += int(b.k != k)
    return rep
