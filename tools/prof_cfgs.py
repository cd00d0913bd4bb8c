import sys, time, cProfile, pstats, io
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch
import slater_oracle as so
import pfaffian_oracle as po
from temfpy_b200 import slater, pfaffian as pf, gutzwiller, engine
be = engine.TorchBackend("cuda:0")
slater._backend = be
H2 = po.bdg_chain(128, t=1.0, mu=0.0, delta=0.05)
C3, _ = so.correlation_matrix(so.hopping_chain(256))
def cfg2(): return pf.H_to_MPS(H2, {"chi_max": 128}, basis="C", _backend=be, as_tenpy=False)
def cfg3a(): return slater.C_to_MPS(C3, {"chi_max": 256}, spinful="PH", as_tenpy=False)
mps3 = cfg3a()
def cfg3b(): return gutzwiller.abrikosov_ph(mps3)
for name, f in (("cfg2 pfaffian", cfg2), ("cfg3 slater part", cfg3a), ("cfg3 gutzwiller part", cfg3b)):
    f()
    t0 = time.perf_counter(); f(); print(name, "%.1f ms" % (1e3 * (time.perf_counter() - t0)))
    pr = cProfile.Profile(); pr.enable(); f(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print("\n".join(s.getvalue().splitlines()[5:26]))
