import os, sys, time
sys.path.insert(0, "/root/repo")
os.environ["TMF_DEBUG_TIMING"] = "1"
import numpy as np
from bench import ground_state_C
from temfpy_b200 import engine
from temfpy_b200.schmidt_utils import to_stopping_condition
be = engine.TorchBackend("cuda:0")
L = 1024
Cm, N = ground_state_C(L)
Cd = be.from_host(Cm.ravel())
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
for it in range(3):
    print("---- iteration", it, file=sys.stderr)
    t0 = time.perf_counter()
    ch = engine.SlaterChain(be, L, tp, N)
    ch.run_modes(Cd, L); t1 = time.perf_counter()
    ch.run_enumerate(); t2 = time.perf_counter()
    ch.run_tensors(Cd, L); t3 = time.perf_counter()
    be.sync(); t4 = time.perf_counter()
    print("py: modes %.1f enumerate %.1f tensors %.1f drain %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)), file=sys.stderr)
    ch.close()
