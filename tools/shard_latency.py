"""Latency of each rank's shard of an N-GPU run, emulated on one GPU (no gather): wall time and kernel timeline."""
import os, sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bench import ground_state_C
from temfpy_b200 import engine, dist as tdist
from temfpy_b200.schmidt_utils import to_stopping_condition
be = engine.TorchBackend("cuda:0")
lib = be.lib
L = 1024
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 1
shards = [int(s) for s in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, world // 2 - 1]
Cm, N = ground_state_C(L)
Cd = be.from_host(Cm.ravel())
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
cuts = tdist.partition(L, world, 1024)
print("partition", cuts)
for s in shards:
    lo, hi = cuts[s]
    for _ in range(3):
        engine.run_chain(be, Cd, L, L, tp, N, site_lo=lo, site_hi=hi, n_chunks=nc, lazy=True).close()
    ws = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r0 = engine.run_chain(be, Cd, L, L, tp, N, site_lo=lo, site_hi=hi, n_chunks=nc, lazy=True)
        torch.cuda.synchronize(); ws.append(1e3 * (time.perf_counter() - t0))
        st = [c.stage_times for c in r0.chains]
        r0.close()
    print("shard", s, (lo, hi), "wall ms", [round(w, 2) for w in ws], "stages", [[round(1e3 * (b - a), 2) for a, b in zip(tt[:-1], tt[1:])] for tt in st])
    lib.tmf_prof_enable(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = engine.run_chain(be, Cd, L, L, tp, N, site_lo=lo, site_hi=hi, n_chunks=nc, lazy=True)
    torch.cuda.synchronize(); print("  profiled wall ms", round(1e3 * (time.perf_counter() - t0), 2))
    buf = C.create_string_buffer(1 << 20)
    lib.tmf_prof_timeline(buf, len(buf))
    lib.tmf_prof_enable(0)
    r.close()
    rows = [ln.split() for ln in buf.value.decode().strip().splitlines()]
    ev = sorted(((r_[0], int(r_[1]), float(r_[2]), float(r_[3])) for r_ in rows), key=lambda e: e[2])
    t_first = ev[0][2]
    prev = None
    for e in ev:
        gap = e[2] - prev if prev is not None else 0.0
        print("   %-14s s%d  %7.3f -> %7.3f  (%.3f)%s" % (e[0], e[1], e[2] - t_first, e[3] - t_first, e[3] - e[2], "   gap %.3f" % gap if gap > 0.02 else ""))
        prev = max(prev or 0, e[3])
