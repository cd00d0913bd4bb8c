"""One conversion of the bench workload (plus one warm-up); target of the ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import ground_state_C
from temfpy_b200 import engine
from temfpy_b200.schmidt_utils import to_stopping_condition
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
be = engine.TorchBackend("cuda:0")
C, N = ground_state_C(L)
Cd = be.from_host(C.ravel())
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
for _ in range(reps):
    res = engine.run_chain(be, Cd, L, L, tp, N, fetch_tensors=False, n_chunks=int(os.environ.get('TMF_RUN_CHUNKS', '0')) or None)
print("ok", res.stats)
