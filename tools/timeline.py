"""GPU timeline of one pipelined conversion (event timestamps per launch and stream)."""
import os, sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bench import ground_state_C
from temfpy_b200 import engine
from temfpy_b200.schmidt_utils import to_stopping_condition
be = engine.TorchBackend("cuda:0")
lib = be.lib
L = 1024
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
Cm, N = ground_state_C(L)
Cd = be.from_host(Cm.ravel())
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
for _ in range(3):
    engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc, lazy=True).close()
os.environ["TMF_DEBUG_TIMING"] = "1"
torch.cuda.synchronize(); t0 = time.perf_counter()
r0 = engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc, lazy=True)
torch.cuda.synchronize(); print("unprofiled wall ms", 1e3 * (time.perf_counter() - t0))
for c in r0.chains:
    tt = c.stage_times
    print("unprofiled chunk", (c.site_lo, c.site_hi), "start %.1f" % (1e3 * (tt[0] - t0)), "stages ms:", [round(1e3 * (b - a), 2) for a, b in zip(tt[:-1], tt[1:])])
r0.close()
os.environ.pop("TMF_DEBUG_TIMING", None)
lib.tmf_prof_enable(1)
torch.cuda.synchronize(); t0 = time.perf_counter()
r = engine.run_chain(be, Cd, L, L, tp, N, n_chunks=nc, lazy=True)
torch.cuda.synchronize(); print("wall ms", 1e3 * (time.perf_counter() - t0))
for c in r.chains:
    tt = c.stage_times
    print("chunk", (c.site_lo, c.site_hi), "start %.1f" % (1e3 * (tt[0] - t0)), "stages ms:", [round(1e3 * (b - a), 2) for a, b in zip(tt[:-1], tt[1:])], "(gate wait, enqueue, modes, enumerate+plan, tensors enqueue, drain)")
buf = C.create_string_buffer(1 << 20)
lib.tmf_prof_timeline(buf, len(buf))
rows = [ln.split() for ln in buf.value.decode().strip().splitlines()]
ev = [(r_[0], int(r_[1]), float(r_[2]), float(r_[3])) for r_ in rows]
end = max(e[3] for e in ev)
print("launches", len(ev), "span ms", end)
# busy time (union of intervals) and per-stream summaries
iv = sorted((e[2], e[3]) for e in ev)
busy, cur0, cur1 = 0.0, iv[0][0], iv[0][1]
for a, b in iv[1:]:
    if a > cur1: busy += cur1 - cur0; cur0, cur1 = a, b
    else: cur1 = max(cur1, b)
busy += cur1 - cur0
print("sum of kernel ms", sum(e[3] - e[2] for e in ev), "union busy ms", busy)
for s in sorted(set(e[1] for e in ev)):
    es = [e for e in ev if e[1] == s]
    print("stream", s, "first", round(min(e[2] for e in es), 2), "last", round(max(e[3] for e in es), 2), "kernel ms", round(sum(e[3] - e[2] for e in es), 2))
    # phases
    ph = {}
    for e in es:
        ph.setdefault(e[0], [1e9, 0, 0.0]); p = ph[e[0]]; p[0] = min(p[0], e[2]); p[1] = max(p[1], e[3]); p[2] += e[3] - e[2]
    print("   ", {k: (round(v[0], 1), round(v[1], 1), round(v[2], 2)) for k, v in sorted(ph.items(), key=lambda kv: kv[1][0])})
