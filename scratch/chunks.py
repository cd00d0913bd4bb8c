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
for nc in [1, 2, 4, 6, 8, 12, 16]:
    ts = []
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc, lazy=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        r.close()
    print("n_chunks", nc, "ms", [round(1e3 * t, 1) for t in ts], flush=True)
