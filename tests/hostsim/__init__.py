"""TEST INFRASTRUCTURE ONLY -- CPU simulator of the CUDA kernels.

``libtemfpy_b200_hostsim.so`` is the very same kernel source (temfpy_b200/csrc/*.cu) compiled by
g++ with -DTMF_HOSTSIM: every CTA program runs sequentially, one simulated thread at a time (see
csrc/cta.hpp).  It lets the CPU suite (`-m "not gpu"`) check kernel *logic* and the host-side
planning without a GPU.  The package never loads it: `temfpy_b200._lib.load()` only accepts the
CUDA build, and the `NumpyBackend` below lives here, not in the package.
"""
import os
import subprocess

import numpy as np

from temfpy_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_PATH = os.environ.get("TMF_SIM_PATH") or os.path.join(HERE, "libtemfpy_b200_hostsim.so")   # (override: ASan build)
CSRC = os.path.join(os.path.dirname(HERE), "..", "temfpy_b200", "csrc")

_sim = None


def load_sim():
    global _sim
    if _sim is None:
        if not os.environ.get("TMF_SIM_PATH"):
            import fcntl
            # several test processes (the two-rank gloo tests) may get here at once: one make at a time
            with open(os.path.join(HERE, ".build.lock"), "w") as lock:
                fcntl.flock(lock, fcntl.LOCK_EX)
                try:
                    subprocess.run(["make", "-s", "-C", CSRC, "hostsim"], check=True, capture_output=True)
                finally:
                    fcntl.flock(lock, fcntl.LOCK_UN)
        _sim = _lib.bind(SIM_PATH)
        assert _sim.tmf_is_cuda() == 0
    return _sim


class NumpyBackend:
    """Host 'device' for the simulator build: buffers are NumPy arrays."""

    def __init__(self):
        self.lib = load_sim()
        self.stream = 0

    def empty(self, n, dtype):
        return np.zeros(max(int(n), 1), dtype=dtype)

    def from_host(self, arr):
        return np.ascontiguousarray(arr).copy()

    def to_host(self, buf, n=None):
        return np.array(buf[:n] if n is not None else buf)

    @staticmethod
    def ptr(buf):
        return buf.ctypes.data

    def sync(self):
        pass

    # pipeline chunks (engine.run_chain with n_chunks > 1): the simulator has no streams
    def side_stream(self, i):
        return None

    def stream_context(self, stream):
        import contextlib
        return contextlib.nullcontext()
