"""Parity of the CUDA path on the *full-size* BASELINE configurations against per-bond fixtures produced
by the reference's own code (oracle/make_golden_full.py): for every bond either bit-exact integer data
(k, filled count, chi, sector table, occupation masks) and Schmidt values within tolerance, or a
documented noise-level ambiguity at the truncation cut (helpers.compare_bonds_fixture)."""
import numpy as np
import pytest

import slater_oracle as so
from tests import helpers

pytestmark = pytest.mark.gpu


def _C_for(name):
    if name == "bonds_cfg1_chain_L64":
        return so.correlation_matrix(so.hopping_chain(64))
    if name == "bonds_cfg5_chain_L1024":
        return so.correlation_matrix(so.hopping_chain(1024))
    if name == "bonds_cfg4_cylinder_6x64":
        return so.correlation_matrix(helpers.cylinder_hamiltonian(64, 6))
    if name == "bonds_cfg3_spinful_ph_L512":
        C1, _ = so.correlation_matrix(so.hopping_chain(256))
        C = so.spinful_correlation_matrix(C1, True)
        return C, int(round(np.trace(C)))
    raise KeyError(name)


@pytest.mark.parametrize("name,min_exact", [("bonds_cfg1_chain_L64", 0.5), ("bonds_cfg3_spinful_ph_L512", 0.5),
                                            ("bonds_cfg4_cylinder_6x64", 0.5), ("bonds_cfg5_chain_L1024", 0.9)])
def test_every_bond_against_reference_fixture(gpu_backend, name, min_exact):
    g = helpers.golden(name)
    C, N = _C_for(name)
    assert N == int(g["N"]) and abs(float(np.sum(C)) - float(g["C_sum"])) < 1e-9     # same input as the reference run
    tp = helpers.golden_trunc(g)
    res = helpers.run_native(gpu_backend, C, tp, N, fetch_tensors=False, **helpers.default_policy(C))
    rep = helpers.compare_bonds_fixture(g, lambda x: res.bonds[x])
    print(f"\n{name}: bonds {rep['bonds']} exact {rep['exact']} ambiguous {rep['ambiguous']} "
          f"max|dchi| {rep['max_dchi']} k_noise {rep['k_noise']} lam_rel {rep['lam_rel']:.2e} "
          f"entropy {rep['entropy']:.2e} options {res.options}")
    assert rep["exact"] + rep["ambiguous"] == rep["bonds"]
    assert rep["lam_rel"] < 1e-12 and rep["entropy"] < 1e-10
    assert rep["exact"] >= min_exact * rep["bonds"], rep
