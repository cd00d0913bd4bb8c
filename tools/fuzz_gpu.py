"""Structured inputs against the oracle on the GPU (robustness sweep)."""
import sys, traceback
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import slater_oracle as so
from tests import helpers
from temfpy_b200.engine import TorchBackend
be = TorchBackend("cuda:0")

def chain(L, t1=-1.0, t2=-1.0, mu=None, pbc=False):
    H = np.zeros((L, L))
    for i in range(L - 1):
        H[i, i + 1] = H[i + 1, i] = t1 if i % 2 == 0 else t2
    if pbc: H[0, L - 1] = H[L - 1, 0] = t2
    if mu is not None: H += np.diag(mu)
    return H

rng = np.random.default_rng(5)
cases = {
    "dimerised strong (t2/t1=0.05)": (chain(160, -1.0, -0.05), {"chi_max": 64}),
    "dimerised moderate (0.5)": (chain(160, -1.0, -0.5), {"chi_max": 64}),
    "weakly coupled halves (1e-6)": (None, {"chi_max": 64}),
    "staggered potential (gapped)": (chain(150, mu=0.8 * (-1.0) ** np.arange(150)), {"chi_max": 64}),
    "random potential (Anderson)": (chain(170, mu=2.0 * rng.standard_normal(170)), {"chi_max": 64}),
    "ring (pbc) + small field": (chain(144, pbc=True, mu=1e-3 * rng.standard_normal(144)), {"chi_max": 96}),
    "svd_min 1e-4": (chain(150, mu=0.1 * rng.standard_normal(150)), {"chi_max": 200, "svd_min": 1e-4}),
    "no chi_max (svd_min 1e-3)": (chain(150, mu=0.1 * rng.standard_normal(150)), {"svd_min": 1e-3}),
    "low filling": (chain(160, mu=1.7 + 0.05 * rng.standard_normal(160)), {"chi_max": 64}),
    "long-range random (decay 3)": (helpers.random_hamiltonian(150, 8), {"chi_max": 64}),
}
Hw = np.zeros((180, 180)); Hw[:90, :90] = helpers.random_hamiltonian(90, 3); Hw[90:, 90:] = helpers.random_hamiltonian(90, 4)
Hw[89, 90] = Hw[90, 89] = 1e-6
cases["weakly coupled halves (1e-6)"] = (Hw, {"chi_max": 64})
bad = 0
for name, (H, tp) in cases.items():
    try:
        Cm, n = so.correlation_matrix(H)
        res = helpers.run_native(be, Cm, tp, n)
        rep = helpers.compare_mps(so.C_to_MPS(Cm, tp), helpers.chain_to_dense(res), tp)
        print(f"OK   {name:34s} N={n:3d} overlap-1={rep.get('overlap', float('nan'))-1:+.1e} entropy={rep['entropy']:.1e} lam_rel={rep['lam_rel']:.1e} ambiguous={rep['ambiguous']}", flush=True)
    except Exception as e:
        bad += 1
        print(f"FAIL {name:34s} {type(e).__name__}: {str(e)[:300]}", flush=True)
print("failures:", bad)
