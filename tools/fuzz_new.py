"""Randomised checks of the round-2 kernels (Gutzwiller projection incl. complex and Pfaffian input, device canonical
form, Pfaffian site finish, centre half modes) against brute force / the oracle.  `python tools/fuzz_new.py sim|gpu N`"""
import itertools, os, sys, time, warnings
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
warnings.simplefilter("ignore")
import numpy as np
import slater_oracle as so
import pfaffian_oracle as po
from tests import helpers
from tests.test_gutzwiller import brute_force, spin_state
from temfpy_b200 import slater, gutzwiller as gw, pfaffian as pf

if sys.argv[1] == "gpu":
    from temfpy_b200 import engine
    be = engine.TorchBackend("cuda:0")
else:
    from tests.hostsim import NumpyBackend
    be = NumpyBackend()
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 0)
fails, done = [], dict(gutz=0, gutz_c=0, pf=0, pf_gutz=0)
t0 = time.time()


def overlap(a, b):
    return abs(np.vdot(a, b)) / (np.linalg.norm(a) * np.linalg.norm(b))


for it in range(N):
    # ---- Slater -> Gutzwiller, real and complex, both conventions, random centre, both canonical forms -------------
    Ls = int(rng.choice([2, 4, 6]))
    cplx = bool(rng.integers(2))
    kind = ["simple", "PH"][int(rng.integers(2))]
    H = helpers.random_hamiltonian(Ls, int(rng.integers(1 << 30)), decay=float(rng.uniform(0.7, 3.0)), cplx=cplx)
    try:
        C_, _ = so.correlation_matrix(H, N=Ls // 2)
        tp = {"chi_max": 4096, "svd_min": 1e-7}
        oc = int(rng.integers(1, 2 * Ls))
        fm = slater.C_to_MPS(C_, tp, spinful=kind, ortho_center=oc, _backend=be, as_tenpy=False)
        total = int(np.asarray(fm.charges[fm.L]).ravel()[0])
        ok_charge = (total == Ls) if kind == "simple" else (total % 2 == 0)
        if ok_charge:
            psi = so.mps_to_state(helpers.block_mps_to_dense(fm))
            phi = brute_force(psi, kind)
            fn = gw.abrikosov if kind == "simple" else gw.abrikosov_ph
            if np.linalg.norm(phi) > 1e-8:
                for mode in (("host", "device") if not cplx else ("host",)):
                    gw.CANONICAL_FORM = mode
                    for canon in (False, True):
                        got = spin_state(fn(fm, return_canonical=canon, _backend=be))
                        o = overlap(phi, got)
                        if not o > 1 - 1e-9:
                            fails.append(("gutz", it, Ls, kind, cplx, oc, mode, canon, o))
                done["gutz_c" if cplx else "gutz"] += 1
    except Exception as e:      # noqa
        fails.append(("gutz-exc", it, Ls, kind, cplx, repr(e)[:200]))
    gw.CANONICAL_FORM = "host"
    # ---- Pfaffian: oracle, device vs host finish, Gutzwiller on the parity-conserving MPS -------------------------
    Lp = int(rng.choice([6, 8, 10, 12, 14]))
    if rng.integers(3) == 0:
        Hb = po.bdg_chain(Lp, mu=float(rng.choice([0.0, 0.3, -0.7])), delta=float(rng.choice([0.05, 0.4, 1.0])))
    else:
        Hb = po.random_bdg(Lp, int(rng.integers(1 << 30)))
    tp = {"chi_max": int(rng.choice([16, 64, 4096])), "svd_min": 1e-7}
    ocp = int(rng.integers(1, Lp)) if rng.integers(2) else None
    try:
        Cm = po.correlation_matrix(Hb, "C->C")
        ref = po.C_to_MPS(Cm, tp, "C", ortho_center=ocp)
        got = pf.C_to_MPS(Cm, tp, basis="C", ortho_center=ocp, _backend=be, as_tenpy=False)
        half = set(range(Lp + 1))
        helpers.compare_pf_mps(ref, helpers.block_mps_to_dense(got), half)
        os.environ["TMF_PF_HOST_FINISH"] = "1"
        goth = pf.C_to_MPS(Cm, tp, basis="C", ortho_center=ocp, _backend=be, as_tenpy=False)
        os.environ.pop("TMF_PF_HOST_FINISH")
        for i in range(Lp):
            d = np.abs(got.get_B_dense(i) - goth.get_B_dense(i)).max()
            if d > 1e-10 * max(np.abs(goth.get_B_dense(i)).max(), 1e-300):
                fails.append(("pf-finish", it, Lp, ocp, i, d))
        done["pf"] += 1
        if Lp <= 10 and tp["chi_max"] == 4096:
            psi = so.mps_to_state(helpers.block_mps_to_dense(got))
            total = int(np.asarray(got.charges[got.L]).ravel()[0])
            for kind, fn in (("simple", gw.abrikosov), ("PH", gw.abrikosov_ph)):
                if total % 2 != ((Lp // 2) % 2 if kind == "simple" else 0):
                    continue
                phi = brute_force(psi, kind)
                if np.linalg.norm(phi) < 1e-8:
                    continue
                o = overlap(phi, spin_state(fn(got, return_canonical=bool(rng.integers(2)), _backend=be)))
                if not o > 1 - 1e-9:
                    fails.append(("pf-gutz", it, Lp, kind, o))
                done["pf_gutz"] += 1
    except NotImplementedError as e:
        pass
    except RuntimeError as e:
        if "zero" not in str(e):
            fails.append(("pf-exc", it, Lp, ocp, tp["chi_max"], repr(e)[:200]))
    except Exception as e:      # noqa
        os.environ.pop("TMF_PF_HOST_FINISH", None)
        fails.append(("pf-exc", it, Lp, ocp, tp["chi_max"], repr(e)[:200]))
print("done", done, "in %.1f s" % (time.time() - t0))
print("FAILS", len(fails))
for f in fails[:20]:
    print("  ", f)
