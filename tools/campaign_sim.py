import sys, time, traceback
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import slater_oracle as so
from tests import helpers
from tests.hostsim import NumpyBackend
be = NumpyBackend()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t_end = time.time() + float(sys.argv[2]) if len(sys.argv) > 2 else time.time() + 300
n_ok = n_bad = 0
while time.time() < t_end:
    L = int(rng.integers(8, 150))
    kind = rng.choice(["random", "anderson", "dimer", "chain_mu", "blocks"])
    if kind == "random":
        H = helpers.random_hamiltonian(L, int(rng.integers(1e6)))
    else:
        H = np.zeros((L, L)); i = np.arange(L - 1)
        H[i, i + 1] = H[i + 1, i] = -1.0
        if kind == "anderson": H += np.diag(rng.uniform(0.5, 6) * rng.standard_normal(L))
        elif kind == "dimer": H[i[1::2], i[1::2] + 1] = H[i[1::2] + 1, i[1::2]] = -rng.uniform(0.01, 1.0)
        elif kind == "chain_mu": H += np.diag(rng.uniform(-1.5, 1.5) + 0.05 * rng.standard_normal(L))
        else:
            c = int(rng.integers(2, L - 2)); H[c - 1, c] = H[c, c - 1] = rng.choice([0.0, 1e-9, 1e-5]); H += np.diag(0.3 * rng.standard_normal(L))
    tp = {"chi_max": int(rng.choice([1, 3, 8, 20, 48, 100]))}
    if rng.random() < 0.3: tp["svd_min"] = float(rng.choice([1e-3, 1e-4, 1e-5, 1e-7]))
    oc = int(rng.integers(1, L)) if rng.random() < 0.4 else None
    try:
        Cm, n = so.correlation_matrix(H)
        if n == 0 or n == L: continue
        try:
            ref = so.C_to_MPS(Cm, tp, ortho_center=oc)
        except Exception as e:
            try:
                helpers.run_native(be, Cm, tp, n, ortho_center=oc)
                print("REF RAISED but ours passed:", kind, L, tp, oc, type(e).__name__, str(e)[:80], flush=True)
            except Exception:
                pass
            continue
        res = helpers.run_native(be, Cm, tp, n, ortho_center=oc)
        helpers.compare_mps(ref, helpers.chain_to_dense(res), tp)
        n_ok += 1
    except Exception as e:
        n_bad += 1
        np.savez(f"/tmp/campaign_fail_{n_bad}.npz", H=H, chi_max=tp["chi_max"], svd_min=tp.get("svd_min", -1.0), oc=-1 if oc is None else oc)
        print("FAIL", kind, "L", L, tp, "oc", oc, type(e).__name__, str(e)[:200], flush=True)
print("ok", n_ok, "bad", n_bad)
