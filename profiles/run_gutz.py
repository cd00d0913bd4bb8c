"""BASELINE configs[2] once: Slater -> MPS of the spinful chain (512 fermion sites, chi = 256), Gutzwiller projection on
the HBM-resident blocks, canonical form on the device (ncu target of profiles/capture_gutz.sh)."""
import sys, warnings
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
warnings.simplefilter("ignore")
import numpy as np, torch
import slater_oracle as so
from temfpy_b200 import slater, gutzwiller, engine
be = engine.TorchBackend("cuda:0"); slater._backend = be
gutzwiller.CANONICAL_FORM = "device"
L = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C3, _ = so.correlation_matrix(so.hopping_chain(L))
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    fm = slater.C_to_MPS(C3, {"chi_max": 256}, spinful="PH", as_tenpy=False)
    sm = gutzwiller.abrikosov_ph(fm)
torch.cuda.synchronize()
print("chi_proj", max(sm.chi), sm.meta["canonical_form"], sm.meta["gemm_jobs"], sm.meta["resident_operands"])
