"""Device time and end-to-end time of one conversion vs chunk order / gate depth / number of chunks."""
import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bench import ground_state_C
from temfpy_b200 import engine, slater
from temfpy_b200.schmidt_utils import to_stopping_condition
be = engine.TorchBackend("cuda:0")
slater._backend = be
L = 1024
Cm, N = ground_state_C(L)
Cd = be.from_host(Cm.ravel())
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
for order in ("natural",):
    os.environ["TMF_CHUNK_ORDER"] = order
    for gate in (3,):
        os.environ.pop("TMF_NO_STAGE_GATE", None); os.environ["TMF_GATE_DEPTH"] = str(gate)
        for nc in (6,):
            for _ in range(2):
                engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc, lazy=True).close()
            ts = []
            for _ in range(9):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                r = engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc, lazy=True)
                torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0)); r.close()
            te = []
            for _ in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                r = engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc)
                torch.cuda.synchronize(); te.append(1e3 * (time.perf_counter() - t0)); del r
            print("order", order, "gate", gate, "n_chunks", nc, "device ms", [round(t, 1) for t in ts], "with D2H+tables ms", [round(t, 1) for t in te], flush=True)
