"""Dumps the per-bond tables of the CUDA path for the full-size fixtures (offline analysis of the parity audit)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
from tests import helpers
from tests.test_gpu_parity_full import _C_for
from temfpy_b200.engine import TorchBackend

be = TorchBackend("cuda:0")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for name in ["bonds_cfg1_chain_L64", "bonds_cfg3_spinful_ph_L512", "bonds_cfg4_cylinder_6x64", "bonds_cfg5_chain_L1024"]:
    g = helpers.golden(name)
    C, N = _C_for(name)
    tp = helpers.golden_trunc(g)
    for snap in (True, False):
        res = helpers.run_native(be, C, tp, N, fetch_tensors=False, snap=snap)
        L = len(C)
        bs = [res.bonds[x] for x in range(L + 1)]
        off = np.concatenate([[0], np.cumsum([b.chi for b in bs])])
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"native_{name}_snap{int(snap)}.npz"),
                            chi_off=off, lam=np.concatenate([b.schmidt_values for b in bs]),
                            charge=np.concatenate([b.charge for b in bs]).astype(np.int32),
                            masks=np.concatenate([b.masks for b in bs]),
                            k=np.array([b.k for b in bs]), fl=np.array([b.filled_left for b in bs]),
                            e=np.array([np.pad(b.e, (0, 64 - len(b.e))) for b in bs]))
        print(name, snap, "done", flush=True)
