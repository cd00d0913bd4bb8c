import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bench import ground_state_C
from temfpy_b200 import engine
from temfpy_b200.schmidt_utils import to_stopping_condition
L = 1024
Cm, N = ground_state_C(L)
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
print(torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else "n/a")
for flat in (0, 1):
    if flat: os.environ["TMF_FLAT_PRIORITY"] = "1"
    else: os.environ.pop("TMF_FLAT_PRIORITY", None)
    be = engine.TorchBackend("cuda:0")
    Cd = be.from_host(Cm.ravel())
    for nc in [3, 4, 6, 8]:
        ts = []
        for it in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc, lazy=True)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            r.close()
        print("flat", flat, "n_chunks", nc, "ms", [round(1e3 * t, 1) for t in ts[1:]], flush=True)
