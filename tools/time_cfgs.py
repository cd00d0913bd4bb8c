"""Wall time of the other BASELINE configurations through the public API (host in, host out)."""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch
import slater_oracle as so
import pfaffian_oracle as po
from tests import helpers
from temfpy_b200 import slater, pfaffian as pf, gutzwiller, engine
be = engine.TorchBackend("cuda:0")
slater._backend = be
def timed(f, n=3):
    f(); ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts), r
C1, _ = so.correlation_matrix(so.hopping_chain(64))
t, m = timed(lambda: slater.C_to_MPS(C1, {"chi_max": 64}, as_tenpy=False)); print("cfg1 chain L=64 chi=64: %.1f ms (%.0f sites/s)" % (1e3 * t, 64 / t))
H2 = po.bdg_chain(128, t=1.0, mu=0.0, delta=0.05)
t, m = timed(lambda: pf.H_to_MPS(H2, {"chi_max": 128}, basis="C", _backend=be, as_tenpy=False)); print("cfg2 Kitaev L=128 chi=128: %.1f ms (%.0f sites/s)" % (1e3 * t, 128 / t))
C3, _ = so.correlation_matrix(so.hopping_chain(256))
def cfg3():
    mps = slater.C_to_MPS(C3, {"chi_max": 256}, spinful="PH", as_tenpy=False)
    return gutzwiller.abrikosov_ph(mps, as_tenpy=False) if "as_tenpy" in gutzwiller.abrikosov_ph.__code__.co_varnames else gutzwiller.abrikosov_ph(mps)
try:
    t, m = timed(cfg3); print("cfg3 Gutzwiller L=256 (512 fermion sites) chi=256: %.1f ms (%.0f spin sites/s)" % (1e3 * t, 256 / t))
except Exception as e:
    print("cfg3 failed:", type(e).__name__, str(e)[:200])
C4, _ = so.correlation_matrix(helpers.cylinder_hamiltonian(64, 6))
t, m = timed(lambda: slater.C_to_MPS(C4, {"chi_max": 1024}, unit_cell_width=64 if False else None, as_tenpy=False)); print("cfg4 cylinder 6x64 chi=1024: %.1f ms (%.0f sites/s)" % (1e3 * t, 384 / t))
# cfg3 split: conversion / projection (device) / canonical form (host sweep)
import warnings
warnings.simplefilter("ignore")
t, fm = timed(lambda: slater.C_to_MPS(C3, {"chi_max": 256}, spinful="PH", as_tenpy=False)); print("cfg3 split: conversion %.1f ms" % (1e3 * t))
t, sm = timed(lambda: gutzwiller.abrikosov_ph(fm, return_canonical=False)); print("cfg3 split: projection (bare tensors to host) %.1f ms, chains %d, resident operands %d" % (1e3 * t, sm.meta["gemm_jobs"], sm.meta["resident_operands"]))
t, sm = timed(lambda: gutzwiller.abrikosov_ph(fm, return_canonical=True)); print("cfg3 split: projection + canonical form %.1f ms, max chi %d" % (1e3 * t, max(sm.chi)))
be.lib.tmf_prof_enable(1)
sm = gutzwiller.abrikosov_ph(fm, return_canonical=True)
import ctypes as C
buf = C.create_string_buffer(1 << 22)
be.lib.tmf_prof_timeline(buf, len(buf))
be.lib.tmf_prof_enable(0)
agg = {}
for ln in buf.value.decode().strip().splitlines():
    t = ln.split(); a = agg.setdefault(t[0], [0, 0.0]); a[0] += 1; a[1] += float(t[3]) - float(t[2])
print("cfg3 canonical kernels:", {k: (v[0], round(v[1], 2)) for k, v in agg.items()}, sm.meta.get("canonical_form"))
