"""BASELINE configs[1] once (BdG p-wave chain L = 128, chi = 128, Pfaffian path) plus the iMPS unit cell of a
trivial-phase chain: ncu target of profiles/capture_pf.sh."""
import sys, warnings
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
warnings.simplefilter("ignore")
import numpy as np, torch
import pfaffian_oracle as po
from temfpy_b200 import pfaffian as pf, engine
be = engine.TorchBackend("cuda:0"); pf._backend = be
H2 = po.bdg_chain(128, t=1.0, mu=0.0, delta=0.05)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    m = pf.H_to_MPS(H2, {"chi_max": 128}, basis="C", _backend=be, as_tenpy=False)
Cs = po.correlation_matrix(po.bdg_chain(64, mu=2.5, delta=0.4), "C->C")
Cl = po.correlation_matrix(po.bdg_chain(66, mu=2.5, delta=0.4), "C->C")
im, err = pf.C_to_iMPS(Cs, Cl, {"chi_max": 96}, 2, 32, basis="C", _backend=be, as_tenpy=False)
torch.cuda.synchronize()
print("chi", max(m.chi), "pfaffians", m.meta["n_pfaffians"], "iMPS chi", [len(l) for l in im.lams], err.left_unitary)
