"""Small conversions through every kernel of the library -- target of compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import slater_oracle as so
import pfaffian_oracle as po
from tests import helpers
from temfpy_b200 import engine, slater, pfaffian, gutzwiller

be = engine.TorchBackend("cuda:0")
slater._backend = be
L = int(os.environ.get("SAN_L", "72"))
C, n = so.correlation_matrix(helpers.random_hamiltonian(L, 11))
for kw in (dict(), dict(nested=False), dict(device_plan=False), dict(snap=True)):
    res = helpers.run_native(be, C, {"chi_max": 32}, n, **kw)
    print("slater", kw, res.stats["path"], flush=True)
Cc, nc = so.correlation_matrix(so.hopping_chain(48))
res = helpers.run_native(be, Cc, {"chi_max": 24, "svd_min": 1e-7}, nc, n_chunks=2)
print("chain 2 chunks", res.stats["max_chi"], flush=True)
m = pfaffian.H_to_MPS(po.bdg_chain(12, mu=0.0, delta=0.05), {"chi_max": 16}, basis="C", as_tenpy=False)
print("pfaffian", m.chi, flush=True)
H1 = np.zeros((24, 24)); H2 = np.zeros((26, 26))
for H in (H1, H2):
    for i in range(len(H) - 1):
        H[i, i + 1] = H[i + 1, i] = -1.0 if i % 2 == 0 else -1.5
im, err = slater.H_to_iMPS(H1, H2, {"chi_max": 30}, 2, 12, as_tenpy=False)
print("iMPS", im.chi, flush=True)
mps = slater.H_to_MPS(so.hopping_chain(8), {"chi_max": 64}, spinful="PH", as_tenpy=False)
sp = gutzwiller.abrikosov_ph(mps)
print("gutzwiller", sp.chi, flush=True)
