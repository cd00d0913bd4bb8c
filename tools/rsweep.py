import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch
import slater_oracle as so
from bench import ground_state_C
from temfpy_b200 import engine
from temfpy_b200.schmidt_utils import to_stopping_condition
be = engine.TorchBackend("cuda:0")
L = 1024
Cm, N = ground_state_C(L)
Cd = be.from_host(Cm.ravel())
tpd = {"chi_max": 1024, "svd_min": 1e-7}
tp = to_stopping_condition(tpd)
trunc = so.Trunc.make(tpd)
ref = {x: so.bond_vectors_from_C(Cm, x, trunc, "LR" if x == 512 else ("R" if x > 512 else "L")) for x in (100, 400, 512, 700)}
for r in (48, 44, 40):
    try:
        res = engine.run_chain(be, Cd, L, L, tp, N, r_sketch=r, fetch_tensors=False)
    except Exception as e:
        print("r", r, "failed", e); continue
    worst = 0.0
    for x, vo in ref.items():
        a = np.sort(vo.lam / np.linalg.norm(vo.lam))[::-1]; b = res.bonds[x].schmidt_values; b = np.sort(b / np.linalg.norm(b))[::-1]
        m = min(len(a), len(b)); worst = max(worst, np.abs(a[:m] / b[:m] - 1).max()); 
        de = abs(so.entropies([a])[0] - so.entropies([b])[0])
        print("  r", r, "bond", x, "chi", len(a), len(b), "max rel dlam %.2e" % np.abs(a[:m] / b[:m] - 1).max(), "dS %.1e" % de)
    ts = []
    for _ in range(7):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rr = engine.run_chain(be, Cd, L, L, tp, N, r_sketch=r, lazy=True)
        torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0)); rr.close()
    print("r", r, "device ms median %.1f min %.1f" % (np.median(ts), min(ts)), flush=True)
