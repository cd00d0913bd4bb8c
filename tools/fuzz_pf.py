"""Robustness sweep of the Pfaffian path against its oracle (GPU)."""
import sys, traceback
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import pfaffian_oracle as po
from tests import helpers
from tests.test_pfaffian import _half_bonds
from temfpy_b200 import pfaffian as pf
from temfpy_b200.engine import TorchBackend
be = TorchBackend("cuda:0")
rng = np.random.default_rng(3)
cases = [
    ("kitaev L=60 mu=1.5 delta=0.3 (trivial)", po.bdg_chain(60, mu=1.5, delta=0.3), {"chi_max": 48}),
    ("kitaev L=61 mu=0.4 delta=0.1", po.bdg_chain(61, mu=0.4, delta=0.1), {"chi_max": 48}),
    ("disordered kitaev L=70", po.bdg_chain(70, mu=0.3, delta=0.2, rng=rng, disorder=1.0), {"chi_max": 48}),
    ("strong disorder L=70", po.bdg_chain(70, mu=0.3, delta=0.2, rng=rng, disorder=4.0), {"chi_max": 32}),
    ("random complex L=24", po.random_bdg(24, 5), {"chi_max": 64}),
    ("random real L=30 decay 1", po.random_bdg(30, 6, decay=1.0, cplx=False), {"chi_max": 64}),
    ("random complex L=20 svd_min 1e-4", po.random_bdg(20, 7), {"chi_max": 200, "svd_min": 1e-4}),
    ("kitaev L=90 delta=0.5", po.bdg_chain(90, mu=0.2, delta=0.5), {"chi_max": 40}),
]
bad = 0
for name, H, tp in cases:
    try:
        Cm = po.correlation_matrix(H, "C->C")
        ref = po.C_to_MPS(Cm, tp, "C")
        got = pf.H_to_MPS(H, tp, basis="C", _backend=be, as_tenpy=False)
        rep = helpers.compare_pf_mps(ref, helpers.block_mps_to_dense(got), _half_bonds(Cm, tp))
        print(f"OK   {name:40s} {rep}", flush=True)
    except Exception as e:
        bad += 1
        print(f"FAIL {name:40s} {type(e).__name__}: {str(e)[:400]}", flush=True)
print("failures:", bad)
