import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bench import ground_state_C
from temfpy_b200 import engine
from temfpy_b200.schmidt_utils import to_stopping_condition
be = engine.TorchBackend("cuda:0")
L = 1024
Cm, N = ground_state_C(L)
Cd = be.from_host(Cm.ravel())
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
for w in ("", "1,1,1,1,1,0.4", "0.5,1,1,1,1,0.4", "1,1,1,1,0.7,0.3", "1,1,1,1,1,0.6,0.25", "0.6,1,1,1,1,1,0.6,0.25"):
    if w: os.environ["TMF_CHUNK_WEIGHTS"] = w
    else: os.environ.pop("TMF_CHUNK_WEIGHTS", None)
    nc = len(w.split(",")) if w else 6
    for _ in range(2):
        engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc, lazy=True).close()
    ts = []
    for _ in range(11):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc, lazy=True)
        torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0)); r.close()
    te = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc)
        torch.cuda.synchronize(); te.append(1e3 * (time.perf_counter() - t0)); del r
    print("weights", w or "equal", "device median %.1f min %.1f" % (np.median(ts), min(ts)), "e2e-ish median %.1f" % np.median(te[1:]), flush=True)
