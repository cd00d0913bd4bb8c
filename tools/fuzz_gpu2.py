"""Second robustness sweep: spinful / ortho_center / sectors / cylinders / odd sizes, against the oracle."""
import sys, traceback
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import slater_oracle as so
from tests import helpers
from temfpy_b200.engine import TorchBackend
be = TorchBackend("cuda:0")
rng = np.random.default_rng(11)
def chain(L, mu=None, t2=-1.0):
    H = np.zeros((L, L))
    for i in range(L - 1): H[i, i + 1] = H[i + 1, i] = -1.0 if i % 2 == 0 else t2
    if mu is not None: H += np.diag(mu)
    return H
cases = []
cases.append(("oc=10 of 140", chain(140, mu=0.05 * rng.standard_normal(140)), {"chi_max": 48}, dict(ortho_center=10)))
cases.append(("oc=130 of 140", chain(140, mu=0.05 * rng.standard_normal(140)), {"chi_max": 48}, dict(ortho_center=130)))
cases.append(("oc=1 of 90", chain(90, mu=0.05 * rng.standard_normal(90)), {"chi_max": 40}, dict(ortho_center=1)))
cases.append(("L=65 (just above direct solver)", chain(65, mu=0.05 * rng.standard_normal(65)), {"chi_max": 64}, {}))
cases.append(("L=129", chain(129, mu=0.05 * rng.standard_normal(129)), {"chi_max": 64}, {}))
cases.append(("L=131 odd, N fixed low", chain(131, mu=1.2 + 0.05 * rng.standard_normal(131)), {"chi_max": 64}, {}))
cases.append(("cylinder 5x20", helpers.cylinder_hamiltonian(20, 5) + 1e-3 * helpers.random_hamiltonian(100, 7), {"chi_max": 200}, {}))
cases.append(("cylinder 3x50", helpers.cylinder_hamiltonian(50, 3) + 1e-3 * helpers.random_hamiltonian(150, 9), {"chi_max": 100}, {}))
cases.append(("chi_max 1 (product)", chain(100, mu=0.05 * rng.standard_normal(100)), {"chi_max": 1}, {}))
cases.append(("chi_max 2", chain(100, mu=0.05 * rng.standard_normal(100)), {"chi_max": 2}, {}))
cases.append(("degeneracy_tol 1e-6", chain(120, mu=0.05 * rng.standard_normal(120)), {"chi_max": 50, "degeneracy_tol": 1e-6}, {}))
cases.append(("strong disorder W=8", chain(200, mu=8.0 * rng.standard_normal(200)), {"chi_max": 32}, {}))
cases.append(("flat band-ish t2=-0.001", chain(160, t2=-0.001), {"chi_max": 32}, {}))
bad = 0
for name, H, tp, kw in cases:
    try:
        Cm, n = so.correlation_matrix(H)
        res = helpers.run_native(be, Cm, tp, n, **kw)
        rep = helpers.compare_mps(so.C_to_MPS(Cm, tp, **kw), helpers.chain_to_dense(res), tp)
        print(f"OK   {name:34s} N={n:3d} overlap-1={rep.get('overlap', float('nan'))-1:+.1e} entropy={rep['entropy']:.1e} ambiguous={rep['ambiguous']}", flush=True)
    except Exception as e:
        bad += 1
        print(f"FAIL {name:34s} {type(e).__name__}: {str(e)[:300]}", flush=True)
print("failures:", bad)
