"""torchrun, N >= 2: dist.C_to_MPS (real and complex input, host segments and NCCL gather) against the single-GPU
slater.C_to_MPS on rank 0."""
import os, sys, warnings
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
warnings.simplefilter("ignore")
import numpy as np, torch, torch.distributed as dist
import slater_oracle as so
from tests import helpers
from temfpy_b200 import engine, slater, dist as tdist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
be = engine.TorchBackend(f"cuda:{local}")
slater._backend = be
ok = True
for L, cplx, chi in ((256, False, 128), (96, True, 64), (512, False, 256)):
    Cm = None
    if rank == 0:
        Cm, _ = so.correlation_matrix(helpers.random_hamiltonian(L, 3, decay=1.5 if cplx else 4.0, cplx=cplx))
    for hx in (True, False):
        mps = tdist.C_to_MPS(Cm, {"chi_max": chi}, backend=be, host_exchange=hx)
        if rank == 0:
            ref = slater.C_to_MPS(Cm, {"chi_max": chi}, as_tenpy=False, _backend=be)
            same = all(np.array_equal(mps.lams[x], ref.lams[x]) and np.array_equal(mps.charges[x], ref.charges[x]) for x in range(L + 1))
            same = same and all(np.array_equal(mps.tensors[i].dense(), ref.tensors[i].dense()) for i in range(0, L, 7))
            print(f"L={L} complex={cplx} host_exchange={hx}: identical to the single-GPU result: {same} "
                  f"({mps.meta['stats']['transport']}, {mps.meta['stats']['n_ranks']} ranks)", flush=True)
            ok = ok and same
if rank == 0:
    print("ALL OK" if ok else "MISMATCH", flush=True)
dist.destroy_process_group()
