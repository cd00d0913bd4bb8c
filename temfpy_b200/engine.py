"""Host-side driver of the native Slater -> MPS chain (``tmf_chain_*`` in the C ABI).

Mirrors the control flow of ``slater.C_to_MPS`` (reference slater.py:1216-1353) but hands the whole
range of sites to the native library at once; Python only allocates the device buffers (PyTorch),
sequences the few native calls and wraps the results.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import SitePlan, check


# ---------------------------------------------------------------------------------------------
# device backend (PyTorch)
# ---------------------------------------------------------------------------------------------
class TorchBackend:
    """Device memory, stream and copies through PyTorch (plumbing only)."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("temfpy_b200 needs a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = _lib.load()
        self.h2d_bytes = 0       # traffic counters (bench.py reports them per step)
        self.d2h_bytes = 0

    def empty(self, n, dtype):
        t = self.torch
        td = {np.float64: t.float64, np.int32: t.int32, np.uint8: t.uint8, np.int64: t.int64}[dtype]
        return t.empty(max(int(n), 1), dtype=td, device=self.device)

    def from_host(self, arr: np.ndarray):
        t = self.torch
        arr = np.ascontiguousarray(arr)
        self.h2d_bytes += arr.nbytes
        return t.from_numpy(arr).to(self.device, non_blocking=False)

    def to_host(self, buf, n=None) -> np.ndarray:
        """Device -> pinned host copy.  The returned array aliases a pinned tensor owned by the
        result (PyTorch's caching host allocator recycles it once the result is dropped)."""
        if n is not None:
            buf = buf[:n]
        t = self.torch
        host = t.empty(buf.shape, dtype=buf.dtype, pin_memory=True)
        host.copy_(buf, non_blocking=True)
        self.sync()
        arr = host.numpy()
        self.d2h_bytes += arr.nbytes
        return arr

    @staticmethod
    def ptr(buf) -> int:
        return buf.data_ptr()

    @property
    def stream(self) -> int:
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def sync(self):
        self.torch.cuda.current_stream(self.device).synchronize()

    # side streams for the chunk pipeline (one per worker thread; PyTorch's current stream is
    # thread-local, so every native call of a worker is enqueued on that worker's stream)
    def side_stream(self, i):
        """Stream of pipeline chunk i.  Earlier chunks get a higher CUDA priority: when the chunks
        compete for SMs in the mode-extraction phase the first one finishes first and its host stage
        (enumeration / planning) runs while the kernels of the later chunks still keep the GPU busy,
        instead of all chunks reaching their host stage at the same moment."""
        if not hasattr(self, "_streams"):
            self._streams = {}
        if i not in self._streams:
            import os
            lo = -5 if not os.environ.get("TMF_FLAT_PRIORITY") else 0
            self._streams[i] = self.torch.cuda.Stream(device=self.device, priority=min(0, lo + i))
        return self._streams[i]

    def stream_context(self, stream):
        return self.torch.cuda.stream(stream)


# ---------------------------------------------------------------------------------------------
# results
# ---------------------------------------------------------------------------------------------
@dataclass
class BondData:
    """Schmidt data of one bond (reference: SchmidtVectors, slater.py:494-543)."""
    x: int
    k: int                      # entangled modes
    filled_left: int
    e: np.ndarray               # (k,) left eigenvalues, decreasing (SchmidtModes.e)
    masks: np.ndarray           # (chi,) uint64: bit i = entangled mode i occupied on the left
    schmidt_values: np.ndarray  # (chi,) un-normalised
    charge: np.ndarray          # (chi,) fermion number to the left
    idx_L: dict                 # charge -> slice

    @property
    def chi(self):
        return len(self.schmidt_values)

    @property
    def sets(self) -> np.ndarray:
        """bool (chi, k) occupation table as in the reference."""
        return ((self.masks[:, None] >> np.arange(self.k, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)


@dataclass
class SiteTensor:
    """Block-sparse site tensor (reference: MPSTensorData.to_npc_array, slater.py:1106-1143)."""
    site: int
    mode: str                   # "left" / "right"
    plan: SitePlan
    blocks: list                # (q_ket, bra_row_start, n_bra_rows, ket_start, n_ket, ndarray[nr, nk])
    row_p: np.ndarray
    row_alpha: np.ndarray
    qtotal: int

    def dense_pab(self) -> np.ndarray:
        """T[p, alpha(bra), beta(ket)] exactly like the oracle's dense_tensor."""
        T = np.zeros((2, self.plan.chi_bra, self.plan.chi_ket))
        for (_, r0, nr, c0, nc, blk) in self.blocks:
            rows = slice(r0, r0 + nr)
            T[self.row_p[rows][:, None], self.row_alpha[rows][:, None], np.arange(c0, c0 + nc)[None, :]] = blk
        return T

    def dense(self) -> np.ndarray:
        """T[vL, p, vR]."""
        T = self.dense_pab()
        return np.transpose(T, (1, 0, 2)) if self.mode == "left" else np.transpose(T, (2, 0, 1))


@dataclass
class ChainResult:
    L: int
    ortho_center: int
    site_lo: int
    site_hi: int
    bonds: dict = field(default_factory=dict)    # x -> BondData
    sites: dict = field(default_factory=dict)    # i -> SiteTensor
    timings: dict = field(default_factory=dict)
    stats: dict = field(default_factory=dict)


# ---------------------------------------------------------------------------------------------
# the driver
# ---------------------------------------------------------------------------------------------
def _ptr_array(p, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)


class SlaterChain:
    """One chain conversion on one device for the sites [site_lo, site_hi)."""

    def __init__(self, backend, L, trunc, n_fermion, ortho_center=None, site_lo=0, site_hi=None,
                 r_sketch=64, n_threads=0):
        self.be = backend
        self.lib = backend.lib
        self.L = int(L)
        self.oc = ortho_center or self.L // 2                      # slater.py:1291
        self.site_lo = int(site_lo)
        self.site_hi = self.L if site_hi is None else int(site_hi)
        sectors = trunc.sector_list(range(0, self.L + 1))
        if sectors is None:
            sec_p, n_sec = None, -1
        else:
            arr = (C.c_int * max(len(sectors), 1))(*sectors)
            sec_p, n_sec = arr, len(sectors)
        chi_max = -1 if trunc.chi_max is None else int(trunc.chi_max)
        self.handle = self.lib.tmf_chain_create(self.L, self.oc, int(n_fermion), chi_max,
                                                float(trunc.svd_min), float(trunc.degeneracy_tol),
                                                sec_p, n_sec, int(r_sketch), self.site_lo,
                                                self.site_hi, int(n_threads))
        if not self.handle:
            raise ValueError(self.lib.tmf_last_error().decode())
        self._buffers = {}

    def close(self):
        if self.handle:
            self.lib.tmf_chain_destroy(self.handle)
            self.handle = None
        self._buffers = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- stage A: modes, enumeration, planning ------------------------------------------------
    def run_modes(self, C_dev, ldc):
        be, lib = self.be, self.lib
        q = (C.c_int64 * 8)()
        check(lib, lib.tmf_chain_modes_sizes(self.handle, q))
        njobs, v_elems, work_bytes = int(q[0]), int(q[1]), int(q[2])
        b = self._buffers
        b["V"] = be.empty(v_elems, np.float64)
        b["e"] = be.empty(njobs * _lib.TMF_MAX_MODES, np.float64)
        b["info"] = be.empty(njobs * 4, np.int32)
        b["work"] = be.empty(work_bytes, np.uint8)
        check(lib, lib.tmf_chain_modes(self.handle, be.ptr(C_dev), int(ldc), be.ptr(b["V"]), be.ptr(b["e"]),
                                       be.ptr(b["info"]), be.ptr(b["work"]), work_bytes, be.stream))
        self.njobs = njobs

    def run_enumerate(self):
        check(self.lib, self.lib.tmf_chain_enumerate(self.handle))

    # -- stage B: tensors -----------------------------------------------------------------------
    def run_tensors(self, C_dev, ldc, out=None):
        be, lib = self.be, self.lib
        q = (C.c_int64 * 8)()
        check(lib, lib.tmf_chain_tensor_sizes(self.handle, q))
        plan_bytes, o_elems, s_elems, nsites, nblocks, out_elems, max_chi = (int(x) for x in q[:7])
        b = self._buffers
        b.pop("work", None)      # the mode workspace is dead by now
        b["plan"] = be.empty(plan_bytes, np.uint8)
        b["O"] = be.empty(o_elems, np.float64)
        b["S"] = be.empty(s_elems, np.float64)
        b["det"] = be.empty(nsites, np.float64)
        b["out"] = out if out is not None else be.empty(out_elems, np.float64)
        self.out_elems, self.nblocks, self.max_chi = out_elems, nblocks, max_chi
        check(lib, lib.tmf_chain_tensors(self.handle, be.ptr(C_dev), int(ldc), be.ptr(b["V"]),
                                         be.ptr(b["plan"]), plan_bytes, be.ptr(b["O"]), be.ptr(b["S"]),
                                         be.ptr(b["det"]), be.ptr(b["out"]), be.stream))

    # -- results ----------------------------------------------------------------------------------
    def bond(self, x) -> BondData:
        lib = self.lib
        q = (C.c_int * 8)()
        lam, charge, masks = _lib.c_double_p(), _lib.c_int_p(), _lib.c_u64_p()
        sec_q, sec_start, e = _lib.c_int_p(), _lib.c_int_p(), _lib.c_double_p()
        check(lib, lib.tmf_chain_bond(self.handle, int(x), q, C.byref(lam), C.byref(charge), C.byref(masks),
                                      C.byref(sec_q), C.byref(sec_start), C.byref(e)))
        chi, k, fl, nsec = int(q[0]), int(q[1]), int(q[2]), int(q[3])
        sq = _ptr_array(sec_q, nsec, np.int64)
        ss = _ptr_array(sec_start, nsec + 1, np.int64)
        idx_L = {int(sq[i]): slice(int(ss[i]), int(ss[i + 1])) for i in range(nsec)}
        return BondData(x=int(x), k=k, filled_left=fl, e=_ptr_array(e, k, np.float64),
                        masks=_ptr_array(masks, chi, np.uint64),
                        schmidt_values=_ptr_array(lam, chi, np.float64),
                        charge=_ptr_array(charge, chi, np.int64), idx_L=idx_L)

    def site(self, i, out_host: np.ndarray) -> SiteTensor:
        lib = self.lib
        plan = SitePlan()
        blocks, boff = _lib.c_int_p(), _lib.c_i64_p()
        row_p, row_a = _lib.c_int_p(), _lib.c_int_p()
        offs = (C.c_int64 * 4)()
        check(lib, lib.tmf_chain_site(self.handle, int(i), C.byref(plan), C.byref(blocks), C.byref(boff),
                                      C.byref(row_p), C.byref(row_a), offs))
        nb = plan.n_blocks
        bl = _ptr_array(blocks, 6 * nb, np.int64).reshape(nb, 6)
        bo = _ptr_array(boff, nb, np.int64)
        out = []
        for b in range(nb):
            r0, nr, c0, nc, _, qk = (int(v) for v in bl[b])
            arr = out_host[bo[b]: bo[b] + nr * nc].reshape(nr, nc)
            out.append((qk, r0, nr, c0, nc, arr))
        return SiteTensor(site=int(i), mode="left" if plan.mode == 0 else "right", plan=plan, blocks=out,
                          row_p=_ptr_array(row_p, plan.n_rows, np.int64),
                          row_alpha=_ptr_array(row_a, plan.n_rows, np.int64), qtotal=plan.qtotal)

    def collect(self, fetch_tensors=True) -> ChainResult:
        """Synchronises and wraps everything into host objects."""
        self.be.sync()
        res = ChainResult(L=self.L, ortho_center=self.oc, site_lo=self.site_lo, site_hi=self.site_hi)
        for x in range(self.site_lo, self.site_hi + 1):
            res.bonds[x] = self.bond(x)
        if fetch_tensors:
            out_host = self.be.to_host(self._buffers["out"], self.out_elems)
            for i in range(self.site_lo, self.site_hi):
                res.sites[i] = self.site(i, out_host)
        res.stats = dict(out_elems=self.out_elems, nblocks=self.nblocks, max_chi=self.max_chi,
                         njobs=self.njobs)
        return res


def _run_range(backend, C_dev, ldc, L, trunc, n_fermion, ortho_center, lo, hi, r_sketch, n_threads,
               fetch_tensors, lazy=False):
    last_err = None
    widths = [r for r in (64, 128, 160) if r > r_sketch]
    for r in [r_sketch] + widths:
        chain = SlaterChain(backend, L, trunc, n_fermion, ortho_center, lo, hi, r, n_threads)
        ok = False
        try:
            chain.run_modes(C_dev, ldc)
            chain.run_enumerate()
            chain.run_tensors(C_dev, ldc)
            if lazy:
                backend.sync()
                ok = True
                return chain
            return chain.collect(fetch_tensors)
        except ValueError as err:           # sketch too narrow -> widen (cylinders)
            last_err = err
            if "r_sketch" not in str(err) or r == 160:
                raise
        finally:
            if not (lazy and ok):
                chain.close()
    raise last_err


class DeviceChainResult:
    """Result of a conversion that stays resident on the device (tensors in ``chain._buffers['out']``
    of every chunk, Schmidt tables in the native chain objects); nothing is wrapped until asked."""

    def __init__(self, chains):
        self.chains = chains

    @property
    def out_elems(self):
        return sum(c.out_elems for c in self.chains)

    def out_buffers(self):
        return [(c._buffers["out"], c.out_elems) for c in self.chains]

    def bond(self, x):
        for c in self.chains:
            if c.site_lo <= x <= c.site_hi:
                return c.bond(x)
        raise KeyError(x)

    def flops(self):
        tot = np.zeros(5)
        for c in self.chains:
            f = (C.c_double * 8)()
            check(c.lib, c.lib.tmf_chain_flops(c.handle, f))
            tot += np.array(f[:5])
        return tot

    def close(self):
        for c in self.chains:
            c.close()
        self.chains = []


def run_chain(backend, C_dev, ldc, L, trunc, n_fermion, ortho_center=None, site_lo=0, site_hi=None,
              r_sketch=64, n_threads=0, fetch_tensors=True, n_chunks=None, lazy=False):
    """C (device) -> Schmidt data of every bond and block-sparse tensor of every site.

    The site range is cut into cost-balanced chunks that run as a software pipeline: one worker
    thread and one CUDA stream per chunk, so that the host stages of a chunk (enumeration,
    planning; the native calls release the GIL) overlap with the kernels of the other chunks.
    Chunks are independent (same decomposition as the multi-GPU shards, see ``dist.py``)."""
    from .dist import partition
    site_hi = L if site_hi is None else site_hi
    nsites = site_hi - site_lo
    if n_chunks is None:
        n_chunks = 4 if (nsites >= 256 and hasattr(backend, "side_stream")) else 1
    n_chunks = max(1, min(n_chunks, nsites))
    if n_chunks == 1:
        r = _run_range(backend, C_dev, ldc, L, trunc, n_fermion, ortho_center, site_lo, site_hi, r_sketch,
                       n_threads, fetch_tensors, lazy)
        return DeviceChainResult([r]) if lazy else r
    cuts = partition(L, n_chunks, trunc.chi_max, ortho_center, lo=site_lo, hi=site_hi)
    backend.sync()                       # C_dev must be complete before the side streams read it

    def work(i):
        lo, hi = cuts[i]
        with backend.stream_context(backend.side_stream(i)):
            return _run_range(backend, C_dev, ldc, L, trunc, n_fermion, ortho_center, lo, hi, r_sketch,
                              n_threads, fetch_tensors, lazy)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=n_chunks) as pool:
        parts = list(pool.map(work, range(n_chunks)))
    if lazy:
        return DeviceChainResult(parts)
    res = ChainResult(L=L, ortho_center=parts[0].ortho_center, site_lo=site_lo, site_hi=site_hi)
    for p in parts:
        res.bonds.update(p.bonds)
        res.sites.update(p.sites)
    res.stats = dict(out_elems=sum(p.stats["out_elems"] for p in parts),
                     nblocks=sum(p.stats["nblocks"] for p in parts),
                     max_chi=max(p.stats["max_chi"] for p in parts),
                     njobs=sum(p.stats["njobs"] for p in parts), n_chunks=n_chunks)
    return res
