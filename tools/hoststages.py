"""Host-side stage timers (TMF_DEBUG_TIMING) of one un-pipelined conversion and of one 128-site shard."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import ground_state_C
from temfpy_b200 import engine
from temfpy_b200.schmidt_utils import to_stopping_condition
be = engine.TorchBackend("cuda:0")
L = 1024
Cm, N = ground_state_C(L)
Cd = be.from_host(Cm.ravel())
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
for lo, hi, thr in ((0, 1024, 0), (448, 576, 2), (448, 576, 0)):
    for _ in range(2):
        engine.run_chain(be, Cd, L, L, tp, N, site_lo=lo, site_hi=hi, n_chunks=1, lazy=True, n_threads=thr).close()
    os.environ["TMF_DEBUG_TIMING"] = "1"
    print(f"===== sites [{lo},{hi}) threads {thr}", file=sys.stderr, flush=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = engine.run_chain(be, Cd, L, L, tp, N, site_lo=lo, site_hi=hi, n_chunks=1, lazy=True, n_threads=thr)
    torch.cuda.synchronize(); print("wall ms %.2f" % (1e3 * (time.perf_counter() - t0)), "stages", [round(1e3 * (b - a), 2) for a, b in zip(r.chains[0].stage_times[:-1], r.chains[0].stage_times[1:])], file=sys.stderr, flush=True)
    r.close()
    os.environ.pop("TMF_DEBUG_TIMING", None)
