"""Device vs host canonical form of Gutzwiller-projected states at two bond dimensions."""
import os, sys, time, warnings
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
warnings.simplefilter("ignore")
import numpy as np, torch
import slater_oracle as so
from temfpy_b200 import slater, gutzwiller, engine
be = engine.TorchBackend("cuda:0"); slater._backend = be
for L, chi in ((256, 256), (128, 1024), (256, 1024)):
    C3, _ = so.correlation_matrix(so.hopping_chain(L))
    fm = slater.C_to_MPS(C3, {"chi_max": chi}, spinful="PH", as_tenpy=False)
    res = {}
    for mode in ("device", "host"):
        os.environ["TMF_CANONICAL_FORM"] = mode
        gutzwiller.abrikosov_ph(fm)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sm = gutzwiller.abrikosov_ph(fm)
        torch.cuda.synchronize(); res[mode] = (time.perf_counter() - t0, sm)
    a, b = res["device"][1], res["host"][1]
    dl = max(np.abs(np.sort(a.lams[x])[::-1][:min(len(a.lams[x]), len(b.lams[x]))] - np.sort(b.lams[x])[::-1][:min(len(a.lams[x]), len(b.lams[x]))]).max() for x in range(a.L + 1))
    print(f"L={L} chi={chi}: projected chi max {max(a.chi)} ({a.meta['canonical_form']}) / {max(b.chi)} ({b.meta['canonical_form']}); "
          f"device {1e3*res['device'][0]:.0f} ms, host {1e3*res['host'][0]:.0f} ms; max |dlam| {dl:.2e}; "
          f"largest sector {max(np.bincount(q - q.min()).max() for q in a.charges if len(q))}", flush=True)
