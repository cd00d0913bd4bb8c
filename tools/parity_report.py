"""Prints the per-bond parity counts of the four full-size fixtures for snap on / off (GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
from tests import helpers
from tests.test_gpu_parity_full import _C_for
from temfpy_b200.engine import TorchBackend

be = TorchBackend("cuda:0")
for name in sys.argv[1:] or ["bonds_cfg1_chain_L64", "bonds_cfg3_spinful_ph_L512", "bonds_cfg4_cylinder_6x64", "bonds_cfg5_chain_L1024"]:
    g = helpers.golden(name)
    C, N = _C_for(name)
    tp = helpers.golden_trunc(g)
    for snap in (True, False):
        for nested in (True, False):
            t = time.time()
            try:
                res = helpers.run_native(be, C, tp, N, fetch_tensors=False, snap=snap, nested=nested)
                rep = helpers.compare_bonds_fixture(g, lambda x: res.bonds[x])
                rep.pop("ambiguous_list")
                print(name, "snap", snap, "nested", nested, rep, res.options, "%.2fs" % (time.time() - t), flush=True)
            except AssertionError as err:
                print(name, "snap", snap, "nested", nested, "AUDIT FAILED:", err, flush=True)
