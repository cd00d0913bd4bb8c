"""Third robustness sweep: spinful conventions, sector filters, unusual truncations (GPU vs oracle)."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import slater_oracle as so
from tests import helpers
from temfpy_b200.engine import TorchBackend
be = TorchBackend("cuda:0")
rng = np.random.default_rng(23)
def chain(L, mu=None):
    H = np.zeros((L, L)); i = np.arange(L - 1); H[i, i + 1] = H[i + 1, i] = -1.0
    if mu is not None: H += np.diag(mu)
    return H
cases = []
H = chain(60, mu=0.3 + 0.05 * rng.standard_normal(60)); C60, _ = so.correlation_matrix(H)
cases.append(("spinful simple L=60 (120 sites)", so.spinful_correlation_matrix(C60, False), {"chi_max": 60}))
cases.append(("spinful PH L=60 (120 sites)", so.spinful_correlation_matrix(C60, True), {"chi_max": 60}))
H = chain(100, mu=0.2 * rng.standard_normal(100)); C100, _ = so.correlation_matrix(H)
cases.append(("spinful PH L=100 chi 200", so.spinful_correlation_matrix(C100, True), {"chi_max": 200}))
Cc, n150 = so.correlation_matrix(chain(150, mu=0.1 * rng.standard_normal(150)))
cases.append(("sectors near half filling", Cc, {"chi_max": 64, "sectors": list(range(0, 151))}))
cases.append(("svd_min 1e-8 (cutoff 1e-16)", Cc, {"chi_max": 100, "svd_min": 1e-8}))
cases.append(("chi_max 3000 svd_min 1e-5", Cc, {"chi_max": 3000, "svd_min": 1e-5}))
bad = 0
for name, Cm, tp in cases:
    try:
        n = int(round(np.trace(Cm)))
        ref = so.C_to_MPS(Cm, tp)
    except Exception as e:
        print(f"REF-RAISES {name:34s} {type(e).__name__}: {str(e)[:120]}")
        try:
            helpers.run_native(be, Cm, tp, int(round(np.trace(Cm))))
            print("     ... but ours succeeded")
        except Exception as e2:
            print(f"     ours raises {type(e2).__name__}: {str(e2)[:120]}")
        continue
    try:
        res = helpers.run_native(be, Cm, tp, n)
        rep = helpers.compare_mps(ref, helpers.chain_to_dense(res), tp)
        print(f"OK   {name:34s} N={n:3d} overlap-1={rep.get('overlap', float('nan'))-1:+.1e} entropy={rep['entropy']:.1e} ambiguous={rep['ambiguous']}", flush=True)
    except Exception as e:
        bad += 1
        print(f"FAIL {name:34s} {type(e).__name__}: {str(e)[:300]}", flush=True)
print("failures:", bad)
