"""Phase times of one multi-GPU step (torchrun)."""
import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch.distributed as dist
from bench import ground_state_C
from temfpy_b200 import engine, dist as tdist
from temfpy_b200.schmidt_utils import to_stopping_condition
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
be = engine.TorchBackend(f"cuda:{local}")
L = 1024
Cm, N = ground_state_C(L)
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
C_dev = be.from_host(Cm.ravel()) if rank == 0 else be.empty(L * L, np.float64)
lo, hi = tdist.partition(L, world, 1024)[rank]
try: ncpu = len(os.sched_getaffinity(0))
except AttributeError: ncpu = os.cpu_count()
nthr = max(2, ncpu // world)
def sync(): torch.cuda.synchronize()
ncs = [int(x) for x in os.environ.get('NCS', '0').split(',')]
for it in range(6 * len(ncs)):
    nc = ncs[it // 6] or None
    dist.barrier(); sync(); t0 = time.perf_counter()
    tdist.broadcast_C(C_dev); sync(); t1 = time.perf_counter()
    res = engine.run_chain(be, C_dev, L, L, tp, N, site_lo=lo, site_hi=hi, n_threads=nthr, lazy=True, n_chunks=nc); sync(); t2 = time.perf_counter()
    bufs = res.out_buffers()
    sync(); t3 = time.perf_counter()
    full, offs = tdist.gather_tensors(bufs); sync(); t4 = time.perf_counter()
    res.close(); dist.barrier(); sync(); t5 = time.perf_counter()
    if it % 6 >= 3:
        print(f"rank {rank} sites [{lo},{hi}) chunks {len(bufs)} thr {nthr}: bcast {1e3*(t1-t0):.2f} chain {1e3*(t2-t1):.2f} cat {1e3*(t3-t2):.2f} gather {1e3*(t4-t3):.2f} close+barrier {1e3*(t5-t4):.2f} total {1e3*(t5-t0):.2f}", flush=True)
dist.destroy_process_group()
