"""Where the end-to-end time of slater.C_to_MPS(C_host) goes (host buffers in, host objects out)."""
import os, sys, time
os.environ['TMF_PY_TIMING'] = '1'
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bench import ground_state_C
from temfpy_b200 import engine, slater
from temfpy_b200.schmidt_utils import to_stopping_condition
be = engine.TorchBackend("cuda:0")
slater._backend = be
L = 1024
C, N = ground_state_C(L)
tpd = {"chi_max": 1024, "svd_min": 1e-7}
tp = to_stopping_condition(tpd)
for _ in range(2):
    slater.C_to_MPS(C, tpd, as_tenpy=False)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Cp = slater._prepare_C(C, None)
    t1 = time.perf_counter()
    Cd = be.from_host(Cp.ravel())
    slater._check_projector(Cp, be=be, Cd=Cd)
    t2 = time.perf_counter()
    res = engine.run_chain(be, Cd, L, L, tp, N)
    t3 = time.perf_counter()
    mps = slater._chain_to_mps(res, L)
    t4 = time.perf_counter()
    print(f"prepare {1e3*(t1-t0):.1f}  h2d {1e3*(t2-t1):.1f}  run_chain {1e3*(t3-t2):.1f}  to_mps {1e3*(t4-t3):.1f}  total {1e3*(t4-t0):.1f} ms")
    for ch in res.timings.get("chunks", []):
        print("    chunk", {k: (round(1e3 * (v - t2), 1) if k == "t_done" else round(v, 1) if k.startswith("d2h_") and k != "d2h_enqueue" else round(1e3 * v, 1)) for k, v in ch.items()})
    del mps, res
