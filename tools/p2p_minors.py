"""torchrun, N >= 2: cost of writing the site tensors straight into rank 0's peer window (dist.FusedGather)
against a local output buffer, per rank: wall time of the shard's chain and the minors kernel time."""
import os, sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch.distributed as dist
from bench import ground_state_C
from temfpy_b200 import engine, dist as tdist
from temfpy_b200.schmidt_utils import to_stopping_condition
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
be = engine.TorchBackend(f"cuda:{local}")
lib = be.lib
L = 1024
Cm, N = ground_state_C(L)
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
C_dev = be.from_host(Cm.ravel())
lo, hi = tdist.partition(L, world, 1024)[rank]
fused = tdist.FusedGather(be)
def kernel_ms(tag):
    buf = C.create_string_buffer(1 << 20)
    lib.tmf_prof_timeline(buf, len(buf))
    return sum(float(r.split()[3]) - float(r.split()[2]) for r in buf.value.decode().strip().splitlines() if r.split()[0] == tag)
for mode in ("local", "peer", "local", "peer"):
    prov = fused if mode == "peer" else None
    ts = []
    for it in range(6):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        res = engine.run_chain(be, C_dev, L, L, tp, N, site_lo=lo, site_hi=hi, lazy=True, n_chunks=1, out_provider=prov)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if prov: fused.complete()
        t2 = time.perf_counter()
        res.close()
        if it >= 2: ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t0)))
    lib.tmf_prof_enable(1)
    dist.barrier(); torch.cuda.synchronize()
    res = engine.run_chain(be, C_dev, L, L, tp, N, site_lo=lo, site_hi=hi, lazy=True, n_chunks=1, out_provider=prov)
    torch.cuda.synchronize()
    if prov: fused.complete()
    km = kernel_ms("minors")
    lib.tmf_prof_enable(0)
    nbytes = 8 * res.out_elems
    res.close()
    print(f"rank {rank} [{lo},{hi}) {mode}: chain ms {np.mean([a for a, _ in ts]):.2f}  +complete {np.mean([b for _, b in ts]):.2f}  minors kernel {km:.3f} ms for {nbytes/1e6:.0f} MB -> {nbytes/1e6/max(km,1e-9):.0f} GB/s", flush=True)
fused.close()
dist.destroy_process_group()
