import sys
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/oracle")
import numpy as np, ctypes as C
from bench import ground_state_C
from temfpy_b200 import engine, _lib
from temfpy_b200.schmidt_utils import to_stopping_condition
be = engine.TorchBackend("cuda:0")
L=1024
Cm,N = ground_state_C(L)
Cd = be.from_host(Cm.ravel())
tp = to_stopping_condition({"chi_max":1024,"svd_min":1e-7})
ch = engine.SlaterChain(be, L, tp, N)
ch.run_modes(Cd, L)
info = be.to_host(ch._buffers["info"]).reshape(-1,4)
sw = info[:,3]
print("sweeps hist", np.bincount(sw[sw>0]))
