import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bench import ground_state_C
from temfpy_b200 import engine
from temfpy_b200.schmidt_utils import to_stopping_condition
be = engine.TorchBackend("cuda:0")
L = 1024
Cm, N = ground_state_C(L)
Cd = be.from_host(Cm.ravel())
tp = to_stopping_condition({"chi_max": 1024, "svd_min": 1e-7})
chain = engine.SlaterChain(be, L, tp, N)
chain.run_modes(Cd, L)
info = chain._buffers["info"].cpu().numpy().reshape(-1, 4)
sw = info[:, 3]
print("jobs", len(sw), "sweeps histogram", np.bincount(sw[sw > 0]))
print("k histogram", np.bincount(info[:, 0]))
chain.close()
