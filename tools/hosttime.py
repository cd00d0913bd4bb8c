import os, sys, time
os.environ["TMF_DEBUG_TIMING"] = "1"
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
lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, L)
for it in range(3):
    print("---- iteration", it, file=sys.stderr)
    chain = engine.SlaterChain(be, L, tp, N, site_lo=lo, site_hi=hi)
    t0 = time.perf_counter(); chain.enqueue_modes(Cd, L); t1 = time.perf_counter()
    chain.finish_modes(); t2 = time.perf_counter()
    chain.run_enumerate(); t3 = time.perf_counter()
    chain.run_tensors(Cd, L); t4 = time.perf_counter()
    be.sync(); t5 = time.perf_counter()
    chain.close()
    print(f"enqueue_modes {1e3*(t1-t0):.2f} finish_modes {1e3*(t2-t1):.2f} enumerate {1e3*(t3-t2):.2f} tensors_enqueue {1e3*(t4-t3):.2f} drain {1e3*(t5-t4):.2f}", file=sys.stderr)
