"""TEST INFRASTRUCTURE ONLY -- import shim for the read-only reference at /root/reference.

Imports the reference's own ``temfpy`` package with the libraries that are absent from this
image (``tenpy``, ``pfapack``) replaced by ``MagicMock`` modules and with a stub
``temfpy._version`` (a hatch-vcs build artefact).  Everything numerical in the reference
(NumPy/LAPACK) then runs unmodified; only the TeNPy packing and the pfapack call are mocked.

Used only by ``oracle/make_golden.py`` (fixture generation, run in the build container where
/root/reference exists) and by CPU tests that are skipped when /root/reference is absent.
Nothing under ``temfpy_b200/`` imports this file.
"""
import importlib.util
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "temfpy"))


def load(test_action: str = "pass"):
    """Returns the reference modules (slater, pfaffian, schmidt_utils, utils, testing)."""
    if not available():
        raise RuntimeError("reference checkout not present at " + REFERENCE_SRC)
    mocks = {
        "tenpy": ["tenpy", "tenpy.linalg", "tenpy.linalg.np_conserved", "tenpy.networks",
                  "tenpy.networks.site", "tenpy.networks.mps"],
        "pfapack": ["pfapack", "pfapack.ctypes"],
    }
    for root, names in mocks.items():
        if importlib.util.find_spec(root) is None:
            for m in names:
                sys.modules.setdefault(m, MagicMock(name=m))
    if "temfpy._version" not in sys.modules:
        v = types.ModuleType("temfpy._version")
        v.__version__ = "0+oracle"
        sys.modules["temfpy._version"] = v
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import temfpy.slater as ref_slater
    import temfpy.pfaffian as ref_pfaffian
    import temfpy.schmidt_utils as ref_su
    import temfpy.utils as ref_utils
    import temfpy.testing as ref_testing
    ref_testing.TEST_ACTION = test_action
    return types.SimpleNamespace(slater=ref_slater, pfaffian=ref_pfaffian, schmidt_utils=ref_su,
                                 utils=ref_utils, testing=ref_testing)
