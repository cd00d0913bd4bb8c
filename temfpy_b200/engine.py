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

    def to_host_async(self, buf, n=None):
        """Enqueues a device -> pinned host copy on the current stream and returns the pinned tensor
        (valid after the stream has been synchronised).  PyTorch's caching host allocator recycles the
        pinned block once the result is dropped."""
        if n is not None:
            buf = buf[:n]
        t = self.torch
        host = t.empty(buf.shape, dtype=buf.dtype, pin_memory=True)
        host.copy_(buf, non_blocking=True)
        self.d2h_bytes += host.numel() * host.element_size()
        return host

    def to_host(self, buf, n=None) -> np.ndarray:
        """Device -> pinned host copy.  The returned array aliases a pinned tensor owned by the
        result."""
        host = self.to_host_async(buf, n)
        self.sync()
        return host.numpy()

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
    dev: tuple | None = None    # (device buffer, element offset of every block) while the shard stays resident in HBM

    def dense_pab(self) -> np.ndarray:
        """T[p, alpha(bra), beta(ket)] exactly like the oracle's dense_tensor."""
        dtype = self.blocks[0][5].dtype if self.blocks else np.float64
        T = np.zeros((2, self.plan.chi_bra, self.plan.chi_ket), dtype=dtype)
        for (_, r0, nr, c0, nc, blk) in self.blocks:
            rows = slice(r0, r0 + nr)
            T[self.row_p[rows][:, None], self.row_alpha[rows][:, None], np.arange(c0, c0 + nc)[None, :]] = blk
        return T

    def dense(self) -> np.ndarray:
        """T[vL, p, vR]."""
        T = self.dense_pab()
        return np.transpose(T, (1, 0, 2)) if self.mode == "left" else np.transpose(T, (2, 0, 1))


class LazyMap:
    """Mapping whose values are wrapped on first access (``factory(key)``) and then cached.  Wrapping
    all 2 L + 1 bond / site objects of a long chain eagerly cost more than the conversion itself."""

    def __init__(self):
        self._made = {}
        self._factory = {}

    def add(self, keys, factory):
        for k in keys:
            self._factory[k] = factory
            self._made.pop(k, None)

    def __getitem__(self, k):
        try:
            return self._made[k]
        except KeyError:
            v = self._made[k] = self._factory[k](k)
            return v

    def __setitem__(self, k, v):
        self._made[k] = v
        self._factory.setdefault(k, None)

    def get(self, k, default=None):
        return self[k] if k in self._factory else default

    def __contains__(self, k):
        return k in self._factory

    def __iter__(self):
        return iter(self._factory)

    def __len__(self):
        return len(self._factory)

    def keys(self):
        return self._factory.keys()

    def values(self):
        return (self[k] for k in self._factory)

    def items(self):
        return ((k, self[k]) for k in self._factory)

    def update(self, other):
        if isinstance(other, LazyMap):
            self._factory.update(other._factory)
            for k in other._factory:
                self._made.pop(k, None)
            self._made.update(other._made)
        else:
            for k, v in dict(other).items():
                self[k] = v


class ShardTables:
    """Host-side results of one shard in bulk arrays (``tmf_chain_bonds_export`` / ``_sites_export``);
    the per-bond and per-site objects are views into them."""

    _STATE = ("first_bond", "n_bonds", "chi_off", "head", "lam", "charge", "masks", "sec_off", "sec_q", "sec_start",
              "e", "site_lo", "blk_off", "blocks", "block_off", "row_off")

    def state(self) -> dict:
        """Picklable form (NumPy arrays + the plan headers as bytes) for the gather of a multi-GPU conversion."""
        st = {k: getattr(self, k) for k in self._STATE if hasattr(self, k)}
        if hasattr(self, "plans"):
            st["plans"] = bytes(self.plans)
        return st

    @classmethod
    def from_state(cls, st: dict, out_host):
        self = cls.__new__(cls)
        for k, v in st.items():
            if k != "plans":
                setattr(self, k, v)
        if "plans" in st:
            n = len(st["plans"]) // C.sizeof(SitePlan)
            self.plans = (SitePlan * max(n, 1)).from_buffer_copy(st["plans"])
        self.out_host = out_host
        return self

    def __init__(self, chain, out_host, want_sites=None):
        lib, h = chain.lib, chain.handle
        q = (C.c_int64 * 4)()
        check(lib, lib.tmf_chain_bonds_sizes(h, q))
        self.first_bond, nb, nchi, nsec = (int(v) for v in q)
        self.n_bonds = nb
        self.chi_off = np.empty(nb + 1, np.int64)
        self.head = np.empty((max(nb, 1), 4), np.int32)
        self.lam = np.empty(nchi, np.float64)
        self.charge = np.empty(nchi, np.int32)
        self.masks = np.empty(nchi, np.uint64)
        self.sec_off = np.empty(nb + 1, np.int64)
        self.sec_q = np.empty(nsec, np.int32)
        self.sec_start = np.empty(nsec + nb, np.int32)
        self.e = np.empty((max(nb, 1), _lib.TMF_MAX_MODES), np.float64)
        p = lambda a: a.ctypes.data
        check(lib, lib.tmf_chain_bonds_export(h, p(self.chi_off), p(self.head), p(self.lam), p(self.charge),
                                              p(self.masks), p(self.sec_off), p(self.sec_q), p(self.sec_start),
                                              p(self.e)))
        self.charge = self.charge.astype(np.int64)
        self.out_host = out_host
        self.site_lo = chain.site_lo
        if out_host is not None or want_sites:
            check(lib, lib.tmf_chain_sites_sizes(h, q))
            ns, nblk, nrows = int(q[0]), int(q[1]), int(q[2])
            self.plans = (SitePlan * max(ns, 1))()
            self.blk_off = np.empty(ns + 1, np.int64)
            self.blocks = np.empty((nblk, 6), np.int32)
            self.block_off = np.empty(nblk, np.int64)
            self.row_off = np.empty(ns + 1, np.int64)
            # the bra-row order is derived lazily per site (see site_rows) instead of exporting 2 chi ints per site
            check(lib, lib.tmf_chain_sites_export(h, C.addressof(self.plans), p(self.blk_off), p(self.blocks),
                                                  p(self.block_off), p(self.row_off), None, None))

    def normalized(self):
        """(norms, normalised Schmidt values) of all bonds of the shard in one vectorised pass; computed once
        (by the chunk's worker thread, next to the device -> host copy of its tensors)."""
        if not hasattr(self, "_lam_n"):
            off = self.chi_off
            lens = np.diff(off)
            valid = lens > 0
            if not valid.any():
                self._norms, self._lam_n = None, None
            else:
                self._norms = np.sqrt(np.add.reduceat(self.lam * self.lam, off[:-1][valid]))
                self._lam_n = self.lam / np.repeat(self._norms, lens[valid])
        return self._norms, self._lam_n

    def bond(self, x) -> "BondData":
        i = x - self.first_bond
        a, b = int(self.chi_off[i]), int(self.chi_off[i + 1])
        s0, s1 = int(self.sec_off[i]), int(self.sec_off[i + 1])
        k, fl = int(self.head[i, 0]), int(self.head[i, 1])
        sq = self.sec_q[s0:s1]
        ss = self.sec_start[s0 + i: s1 + i + 1]
        idx_L = {int(sq[j]): slice(int(ss[j]), int(ss[j + 1])) for j in range(s1 - s0)}
        return BondData(x=int(x), k=k, filled_left=fl, e=self.e[i, :k], masks=self.masks[a:b],
                        schmidt_values=self.lam[a:b], charge=self.charge[a:b], idx_L=idx_L)

    def site_rows(self, i, plan):
        """(row_p, row_alpha) of site i: bra rows in the reference order -- [p = 0 | p = 1], stably sorted by the
        pipe charge q(alpha) + p (left tensors) / q(alpha) - p (right tensors), slater.py:1053-1058."""
        x = i if plan.mode == 0 else i + 1                    # bra bond
        k = x - self.first_bond
        a, b = int(self.chi_off[k]), int(self.chi_off[k + 1])
        q = self.charge[a:b]
        chi = b - a
        p = np.repeat(np.arange(2, dtype=np.int64), chi)
        al = np.tile(np.arange(chi, dtype=np.int64), 2)
        order = np.argsort(q[al] + (p if plan.mode == 0 else -p), kind="stable")
        return p[order], al[order]

    def site(self, i) -> "SiteTensor":
        u = i - self.site_lo
        plan = self.plans[u]
        b0, b1 = int(self.blk_off[u]), int(self.blk_off[u + 1])
        out = []
        for b in range(b0, b1):
            r0, nr, c0, nc, _, qk = (int(v) for v in self.blocks[b])
            o = int(self.block_off[b])
            out.append((qk, r0, nr, c0, nc, self.out_host[o: o + nr * nc].reshape(nr, nc)))
        row_p, row_alpha = self.site_rows(i, plan)
        dev = None
        if getattr(self, "out_dev", None) is not None:
            dev = (self.out_dev, [int(v) for v in self.block_off[b0:b1]])
        return SiteTensor(site=int(i), mode="left" if plan.mode == 0 else "right", plan=plan, blocks=out,
                          row_p=row_p, row_alpha=row_alpha, qtotal=plan.qtotal, dev=dev)


@dataclass
class ChainResult:
    L: int
    ortho_center: int
    site_lo: int
    site_hi: int
    bonds: LazyMap = field(default_factory=LazyMap)    # x -> BondData   (wrapped on first access)
    sites: LazyMap = field(default_factory=LazyMap)    # i -> SiteTensor
    tables: list = field(default_factory=list)         # ShardTables of every chunk
    timings: dict = field(default_factory=dict)
    stats: dict = field(default_factory=dict)
    options: dict = field(default_factory=dict)        # r_sketch / snap / nested the conversion ended up with

    def lam_charge(self, x):
        """(Schmidt values, charges) of bond x without wrapping a BondData."""
        for t in self.tables:
            i = x - t.first_bond
            if 0 <= i < t.n_bonds and t.chi_off[i + 1] > t.chi_off[i]:
                a, b = int(t.chi_off[i]), int(t.chi_off[i + 1])
                return t.lam[a:b], t.charge[a:b]
        b = self.bonds[x]
        return b.schmidt_values, b.charge


def bulk_normalized(res: "ChainResult", L: int, log=None):
    """Normalised Schmidt vectors (slater.py:1296, utils.normalize_SV) and charges of all L + 1 bonds from
    the shard tables: one vectorised pass per shard instead of 2 L small NumPy calls."""
    lams, charges = [None] * (L + 1), [None] * (L + 1)
    for t in res.tables:
        norms, lam_n = t.normalized()
        if lam_n is None:
            continue
        if log is not None and log.isEnabledFor(20):
            for nv in norms:
                log.info("Norm of Schmidt values: %s", nv)
        off = t.chi_off
        for i in np.flatnonzero(np.diff(off) > 0):
            x = t.first_bond + int(i)
            if 0 <= x <= L:
                a, b = int(off[i]), int(off[i + 1])
                lams[x] = lam_n[a:b]
                charges[x] = t.charge[a:b]
    return lams, charges


class LazySeq:
    """Read-only sequence view of a LazyMap over range(n) (the site tensors of an MPS)."""

    def __init__(self, mapping, n):
        self._m, self._n = mapping, n

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._m[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._m[i]

    def __iter__(self):
        return (self._m[i] for i in range(self._n))


# ---------------------------------------------------------------------------------------------
# the driver
# ---------------------------------------------------------------------------------------------
def _ptr_array(p, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)


_REAPER = None


def _reaper():
    """Queue of (lib, chain handle) pairs destroyed by a daemon thread (ctypes releases the GIL)."""
    global _REAPER
    if _REAPER is None:
        import queue
        import threading
        q = queue.SimpleQueue()

        def run():
            while True:
                lib, h = q.get()
                try:
                    lib.tmf_chain_destroy(h)
                except Exception:       # pragma: no cover - interpreter shutdown
                    pass
        threading.Thread(target=run, daemon=True, name="tmf-chain-reaper").start()
        _REAPER = q
    return _REAPER


class SlaterChain:
    """One chain conversion on one device for the sites [site_lo, site_hi)."""

    def __init__(self, backend, L, trunc, n_fermion, ortho_center=None, site_lo=0, site_hi=None,
                 r_sketch=48, n_threads=0, snap=False, nested=None, device_plan=None, cplx=False):
        self.be = backend
        self.cplx = bool(cplx)
        self.es = 2 if cplx else 1       # doubles per tensor element
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
        if not snap:
            check(self.lib, self.lib.tmf_chain_set_option(self.handle, _lib.OPT_SNAP, 0))
        if nested is not None:      # None: the library's default (nested unless TMF_LEGACY_FILLED is set)
            check(self.lib, self.lib.tmf_chain_set_option(self.handle, _lib.OPT_NESTED, int(bool(nested))))
        if device_plan is not None:
            check(self.lib, self.lib.tmf_chain_set_option(self.handle, _lib.OPT_DEVICE_PLAN, int(bool(device_plan))))
        if cplx:      # C_dev of the stages below is the 2L x 2L real embedding, r_sketch counts real columns
            check(self.lib, self.lib.tmf_chain_set_option(self.handle, _lib.OPT_COMPLEX, 1))
        self._buffers = {}

    def close(self):
        """Releases the device buffers now and the native chain object (tens of MB of host tables) on a
        background thread: freeing them costs ~1 ms that nobody has to wait for."""
        if self.handle:
            _reaper().put((self.lib, self.handle))
            self.handle = None
        self._buffers = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- stage A: modes, enumeration, planning ------------------------------------------------
    def enqueue_modes(self, C_dev, ldc):
        """Enqueues the mode extraction of every (bond, side) of the shard; does not wait."""
        be, lib = self.be, self.lib
        q = (C.c_int64 * 8)()
        check(lib, lib.tmf_chain_modes_sizes(self.handle, q))
        njobs, v_elems, work_bytes, e_elems = int(q[0]), int(q[1]), int(q[2]), int(q[3])
        b = self._buffers
        b["V"] = be.empty(v_elems, np.float64)
        b["e"] = be.empty(max(e_elems, njobs * _lib.TMF_MAX_MODES), np.float64)
        b["info"] = be.empty(njobs * 4, np.int32)
        b["work"] = be.empty(work_bytes, np.uint8)
        check(lib, lib.tmf_chain_modes_enqueue(self.handle, be.ptr(C_dev), int(ldc), be.ptr(b["V"]),
                                               be.ptr(b["e"]), be.ptr(b["info"]), be.ptr(b["work"]),
                                               work_bytes, be.stream))
        self.njobs = njobs

    def finish_modes(self):
        """Spectra to the host (waits for the stream)."""
        b = self._buffers
        check(self.lib, self.lib.tmf_chain_modes_finish(self.handle, self.be.ptr(b["e"]), self.be.ptr(b["info"]),
                                                        self.be.stream))

    def run_modes(self, C_dev, ldc):
        self.enqueue_modes(C_dev, ldc)
        self.finish_modes()

    def run_enumerate(self):
        """Subset enumeration on the device (one warp per bond), planning on the host."""
        import os
        if os.environ.get("TMF_HOST_ENUMERATE"):
            check(self.lib, self.lib.tmf_chain_enumerate(self.handle))
            return
        be, lib = self.be, self.lib
        nbytes = int(lib.tmf_chain_enum_workspace(self.handle))
        work = self._buffers.get("work")
        if work is None or work.nbytes < nbytes:
            work = self._buffers["work"] = be.empty(nbytes, np.uint8)     # the mode workspace is dead by now
        check(lib, lib.tmf_chain_enumerate_dev(self.handle, be.ptr(work), nbytes, be.stream))

    # -- stage B: tensors -----------------------------------------------------------------------
    def tensor_doubles(self) -> int:
        """Doubles of the shard's site tensors (known once the bonds are enumerated)."""
        q = (C.c_int64 * 8)()
        check(self.lib, self.lib.tmf_chain_tensor_sizes(self.handle, q))
        return self.es * int(q[5])

    def run_tensors(self, C_dev, ldc, out=None):
        be, lib = self.be, self.lib
        q = (C.c_int64 * 8)()
        check(lib, lib.tmf_chain_tensor_sizes(self.handle, q))
        plan_bytes, o_elems, s_elems, nsites, nblocks, out_elems, max_chi = (int(x) for x in q[:7])
        self.path = dict(nested=bool(int(q[7]) & 1), device_plan=bool(int(q[7]) & 2))
        b = self._buffers           # ("work" stays: it holds the resident enumeration tables / site plans)
        es = self.es
        b["plan"] = be.empty(plan_bytes, np.uint8)
        b["O"] = be.empty(es * o_elems, np.float64)
        b["S"] = be.empty(es * s_elems, np.float64)
        b["det"] = be.empty(es * nsites, np.float64)
        b["out"] = out if out is not None else be.empty(es * out_elems, np.float64)
        self.out_elems, self.nblocks, self.max_chi = out_elems, nblocks, max_chi
        check(lib, lib.tmf_chain_tensors(self.handle, be.ptr(C_dev), int(ldc), be.ptr(b["V"]),
                                         be.ptr(b["plan"]), plan_bytes, be.ptr(b["O"]), be.ptr(b["S"]),
                                         be.ptr(b["det"]), be.ptr(b["out"]), be.stream))
        # the per-site determinants (8 bytes each) come back with a plain copy: a NaN among them reports a broken
        # elimination (check_det) -- no library reduction kernels on the stream
        if hasattr(be, "to_host_async"):
            b["det_host"] = be.to_host_async(b["det"], es * nsites)

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

    def check_det(self):
        """The site kernels report an elimination that broke down through a NaN determinant: filled spaces that
        are not nested (nested kernel) or always-occupied orbitals of the two bonds that are orthogonal (Schur
        kernel: incompatible truncations of a degenerate multiplet).  The driver then redoes the conversion with
        other options.  Needs the stream to be synchronised."""
        det = self._buffers.get("det_host")
        if det is None:
            d = self._buffers.get("det")
            if d is None:
                return
            det = self.be.to_host(d) if hasattr(self.be, "torch") else np.asarray(d)
        else:
            det = det.numpy() if hasattr(det, "numpy") else np.asarray(det)
        ok = bool(np.all(np.isfinite(det)))
        if not ok:
            raise ValueError("site stage: singular elimination (incompatible neighbouring bonds)")

    def collect(self, fetch_tensors=True) -> ChainResult:
        """Brings the tensors to the host (one pinned copy) and exports the host-side tables in bulk;
        the per-bond / per-site objects are wrapped lazily from those."""
        import time
        t0 = time.perf_counter()
        res = ChainResult(L=self.L, ortho_center=self.oc, site_lo=self.site_lo, site_hi=self.site_hi)
        host = None
        evs = None
        if fetch_tensors and hasattr(self.be, "to_host_async"):
            import os
            if os.environ.get("TMF_PY_TIMING"):
                tc = self.be.torch.cuda
                evs = (tc.Event(enable_timing=True), tc.Event(enable_timing=True))
                evs[0].record(tc.current_stream(self.be.device))
            host = self.be.to_host_async(self._buffers["out"], self.es * self.out_elems)   # overlaps the table export
            if evs:
                evs[1].record(self.be.torch.cuda.current_stream(self.be.device))
        t1 = time.perf_counter()
        if fetch_tensors and host is None:
            out_host = self.be.to_host(self._buffers["out"], self.es * self.out_elems)
        else:
            out_host = host.numpy() if host is not None else None
        if out_host is not None and self.cplx:
            out_host = out_host.view(np.complex128)
        tab = ShardTables(self, out_host)
        tab._pinned = host
        if getattr(self, "keep_device", False) and fetch_tensors:
            tab.out_dev = self._buffers["out"]        # the shard's tensors stay addressable in HBM (Gutzwiller)
        tab.normalized()
        t2 = time.perf_counter()
        self.be.sync()
        self.check_det()
        t3 = time.perf_counter()
        res.tables = [tab]
        res.bonds.add([x for x in range(tab.first_bond, tab.first_bond + tab.n_bonds)
                       if self.site_lo <= x <= self.site_hi or x == self.oc], tab.bond)
        if fetch_tensors:
            res.sites.add(range(self.site_lo, self.site_hi), tab.site)
        res.timings = dict(d2h_enqueue=t1 - t0, tables=t2 - t1, sync=t3 - t2)
        if evs:
            res.timings["d2h_ms"] = evs[0].elapsed_time(evs[1])
            res.timings["d2h_GBps"] = 8e-6 * self.out_elems / max(res.timings["d2h_ms"], 1e-9)
            res.timings["t_done"] = time.perf_counter()
        res.stats = dict(out_elems=self.out_elems, nblocks=self.nblocks, max_chi=self.max_chi,
                         njobs=self.njobs, path=self.path)
        return res


class _NoGate:
    def before(self):
        pass

    def after(self):
        pass


class _Retry(Exception):
    """A chunk asks for the whole conversion to be redone with other options (kept global so that every
    chunk / rank builds a shared boundary bond with the same kernels and sketch width)."""

    def __init__(self, kind, err):
        super().__init__(str(err))
        self.kind, self.err = kind, err


SKETCH_WIDTHS = (48, 64, 128, 160)


def _run_range(backend, C_dev, ldc, L, trunc, n_fermion, ortho_center, lo, hi, r_sketch, n_threads,
               fetch_tensors, lazy=False, gate=None, snap=False, nested=None, device_plan=None, cplx=False,
               keep_device=False, out_provider=None):
    import time
    gate = gate or _NoGate()
    provided = False
    chain = SlaterChain(backend, L, trunc, n_fermion, ortho_center, lo, hi, r_sketch, n_threads, snap=snap,
                        nested=nested, device_plan=device_plan, cplx=cplx)
    chain.keep_device = keep_device
    ok = False
    try:
        tt = [time.perf_counter()]
        try:
            gate.before()
            tt.append(time.perf_counter())
            chain.enqueue_modes(C_dev, ldc)
        finally:
            gate.after()
        tt.append(time.perf_counter())
        chain.finish_modes()
        tt.append(time.perf_counter())
        chain.run_enumerate()
        tt.append(time.perf_counter())
        out = None
        if out_provider is not None:      # (multi-GPU: a slice of the destination rank's peer window, dist.py)
            provided = True
            out = out_provider(chain)
            if not hasattr(out, "device") and hasattr(backend, "torch"):     # memory of another GPU
                check(chain.lib, chain.lib.tmf_chain_set_option(chain.handle, _lib.OPT_PEER_OUT, 1))
        chain.run_tensors(C_dev, ldc, out=out)
        tt.append(time.perf_counter())
        if lazy:
            backend.sync()
            chain.check_det()
            tt.append(time.perf_counter())
            chain.stage_times = tt      # gate wait, enqueue, modes, enumerate + plan, tensors enqueue, drain
            ok = True
            return chain
        return chain.collect(fetch_tensors)
    except ValueError as err:
        msg = str(err)
        if "r_sketch" in msg:               # sketch too narrow -> widen (cylinders)
            raise _Retry("sketch", err) from None
        if "nested site stage" in msg:      # filled spaces not nested within the threshold noise -> explicit bases
            raise _Retry("nested", err) from None
        if "singular elimination" in msg:
            raise _Retry("singular", err) from None
        raise
    finally:
        if out_provider is not None and not provided:
            out_provider.abort()          # the other ranks must not wait for this one's size
        if not (lazy and ok):
            chain.close()


class _StageGate:
    """Orders the mode-extraction stages of consecutive pipeline chunks on the device: chunk i's first
    kernel waits for chunk i-1's last mode kernel (an event; the host never blocks on the GPU).  Run
    all at once, the stages share the SMs, finish together and leave the GPU idle while every chunk is in
    its host stage; run one after the other, the host stage and the tensor kernels of chunk i overlap the
    mode kernels of chunk i+1."""

    def __init__(self, backend, n):
        import threading
        self.be = backend
        import os
        self.enqueued = [threading.Event() for _ in range(n)]
        self.done = [None] * n
        self.depth = max(1, int(os.environ.get("TMF_GATE_DEPTH", "3")))

    def gate(self, i):
        outer = self

        class G:
            used = False

            def before(self):
                j = i - outer.depth        # `depth` mode stages may be in flight at once
                if self.used or j < 0:
                    return
                outer.enqueued[j].wait()
                ev = outer.done[j]
                if ev is not None:
                    outer.be.torch.cuda.current_stream(outer.be.device).wait_event(ev)

            def after(self):
                if self.used:
                    return
                self.used = True
                try:
                    ev = outer.be.torch.cuda.Event()
                    ev.record(outer.be.torch.cuda.current_stream(outer.be.device))
                    outer.done[i] = ev
                finally:
                    outer.enqueued[i].set()
        return G()


class DeviceChainResult:
    """Result of a conversion that stays resident on the device (tensors in ``chain._buffers['out']``
    of every chunk, Schmidt tables in the native chain objects); nothing is wrapped until asked."""

    def __init__(self, chains):
        self.chains = chains

    @property
    def out_elems(self):
        return sum(c.out_elems for c in self.chains)

    def out_buffers(self):
        return [(c._buffers["out"], c.es * c.out_elems) for c in self.chains]

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
              r_sketch=48, n_threads=0, fetch_tensors=True, n_chunks=None, lazy=False, snap=False, nested=None,
              device_plan=None, cplx=False, keep_device=False, out_provider=None):
    """C (device) -> Schmidt data of every bond and block-sparse tensor of every site.

    The site range is cut into cost-balanced chunks that run as a software pipeline: one worker
    thread and one CUDA stream per chunk, so that the host stages of a chunk (enumeration,
    planning; the native calls release the GIL) overlap with the kernels of the other chunks.
    Chunks are independent (same decomposition as the multi-GPU shards, see ``dist.py``).

    Three conditions make the *whole* conversion start over with other options, for all chunks alike (a
    boundary bond is computed by both neighbouring chunks and must come out of identical kernels): a range
    sketch that is too narrow for the entanglement spectrum (widened 48 -> 64 -> 128 -> 160; cylinders),
    neighbouring bonds whose kept Schmidt vectors are incompatible (``snap``: the truncation then never cuts
    inside a numerically degenerate multiplet) and filled spaces that are not nested within the threshold
    noise (explicit filled bases instead).

    ``snap=False`` (default) is the reference's literal truncation (schmidt_utils.py:140-185 only sees
    degeneracies below ``degeneracy_tol``; where a multiplet that is degenerate in exact arithmetic straddles
    ``chi_max`` the kept part is decided by the rounding noise of the mode eigenvalues, in the reference as here)."""
    opts = dict(r_sketch=r_sketch, snap=snap, nested=default_nested(nested, trunc, cplx), device_plan=device_plan,
                cplx=cplx, keep_device=keep_device, out_provider=out_provider)
    while True:
        try:
            return _run_chain_once(backend, C_dev, ldc, L, trunc, n_fermion, ortho_center, site_lo, site_hi,
                                   n_threads, fetch_tensors, n_chunks, lazy, opts)
        except _Retry as rt:
            if rt.kind == "peer":           # multi-rank callers decide for all ranks together (dist.C_to_MPS)
                raise
            if rt.kind == "sketch":
                wider = [r for r in SKETCH_WIDTHS if r > opts["r_sketch"]]
                if not wider:
                    raise rt.err
                opts["r_sketch"] = wider[0]
            elif rt.kind == "singular" and not opts["snap"]:
                opts["snap"] = True             # never cut inside a numerically degenerate multiplet
            elif opts["nested"] is False or opts.get("cplx"):
                raise rt.err
            else:
                opts["nested"] = False


NESTED_MAX_SVD_MIN = 3e-5


def default_nested(nested, trunc, cplx=False):
    """The nested-projector site stage is exact up to O(cutoff), cutoff = svd_min^2: modes below the cutoff count as
    exactly filled / empty when the filled spaces of neighbouring bonds are related.  At the default svd_min = 1e-6
    that is 1e-12; for coarse truncations (svd_min = 1e-3: overlap with the reference 1 - 4e-9, found by
    tools/campaign_sim.py) the explicit filled bases are used instead (exact; the coarse conversions are cheap).
    Complex determinants have the nested form only."""
    if nested is None and not cplx and trunc.svd_min > NESTED_MAX_SVD_MIN:
        return False
    return nested


def _public_opts(opts):
    return {k: v for k, v in opts.items() if k != "out_provider"}


def _run_chain_once(backend, C_dev, ldc, L, trunc, n_fermion, ortho_center, site_lo, site_hi, n_threads,
                    fetch_tensors, n_chunks, lazy, opts):
    from .dist import partition
    r_sketch, snap, nested, device_plan = opts["r_sketch"], opts["snap"], opts["nested"], opts["device_plan"]
    cplx = opts.get("cplx", False)
    keep = opts.get("keep_device", False)
    provider = opts.get("out_provider")
    if provider is not None and not getattr(provider, "multi_chunk", False):
        n_chunks = 1                     # one slice of the peer window per rank
    site_hi = L if site_hi is None else site_hi
    nsites = site_hi - site_lo
    if n_chunks is None:
        # (measured on B200 at L = 1024: 3 chunks 14.8 ms, 6 chunks 15.6 ms, 1 chunk 18 ms per conversion -- the host
        #  stages are short since the site planning moved to the device, more chunks only add contention)
        n_chunks = (3 if nsites >= 384 else 2 if nsites >= 192 else 1) if hasattr(backend, "torch") else 1
    n_chunks = max(1, min(n_chunks, nsites))
    if n_chunks == 1:
        r = _run_range(backend, C_dev, ldc, L, trunc, n_fermion, ortho_center, site_lo, site_hi, r_sketch,
                       n_threads, fetch_tensors, lazy, snap=snap, nested=nested, device_plan=device_plan, cplx=cplx,
                       keep_device=keep, out_provider=provider)
        if lazy:
            r = DeviceChainResult([r])
        r.options = _public_opts(opts)
        return r
    import os
    # equal-cost chunks with a short last one: after the last mode stage only its host stage and tensor
    # kernels remain, and those scale with its number of sites
    weights = [1.0] * (n_chunks - 1) + [0.4] if n_chunks >= 4 else None
    if os.environ.get("TMF_CHUNK_WEIGHTS"):
        weights = [float(v) for v in os.environ["TMF_CHUNK_WEIGHTS"].split(",")]
    cuts = partition(L, n_chunks, trunc.chi_max, ortho_center, lo=site_lo, hi=site_hi, weights=weights)
    backend.sync()                       # C_dev must be complete before the side streams read it

    stages = _StageGate(backend, n_chunks) if hasattr(backend, "torch") and not os.environ.get("TMF_NO_STAGE_GATE") \
        else None

    # pipeline order: natural (left to right) unless TMF_CHUNK_ORDER=ends asks for "chain ends first, centre
    # last"; measured within noise of each other on B200, the natural order brings the tensors to the host sooner
    order = list(range(n_chunks))
    if os.environ.get("TMF_CHUNK_ORDER") == "ends":
        oc_ = ortho_center or L // 2
        order = sorted(range(n_chunks), key=lambda i: -abs(0.5 * (cuts[i][0] + cuts[i][1]) - oc_))

    def work(pos):
        lo, hi = cuts[order[pos]]
        with backend.stream_context(backend.side_stream(pos)):
            try:
                return _run_range(backend, C_dev, ldc, L, trunc, n_fermion, ortho_center, lo, hi, r_sketch,
                                  n_threads, fetch_tensors, lazy, gate=stages.gate(pos) if stages else None,
                                  snap=snap, nested=nested, device_plan=device_plan, cplx=cplx, keep_device=keep,
                                  out_provider=provider)
            except _Retry as rt:        # the other chunks finish; the driver then starts over
                return rt

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=n_chunks) as pool:
        done = list(pool.map(work, range(n_chunks)))
    retries = [r for r in done if isinstance(r, _Retry)]
    if retries:
        for r in done:
            if lazy and not isinstance(r, _Retry):
                r.close()
        # a too narrow sketch takes precedence: the nested check is only meaningful on complete spectra
        raise next((r for r in retries if r.kind == "sketch"), retries[0])
    parts = [None] * n_chunks
    for pos, r in enumerate(done):
        parts[order[pos]] = r
    if lazy:
        out = DeviceChainResult(parts)
        out.options = _public_opts(opts)
        return out
    res = ChainResult(L=L, ortho_center=parts[0].ortho_center, site_lo=site_lo, site_hi=site_hi)
    for p in parts:
        res.bonds.update(p.bonds)
        res.sites.update(p.sites)
    res.timings = dict(chunks=[p.timings for p in parts])
    res.tables = [t for p in parts for t in p.tables]
    res.stats = dict(out_elems=sum(p.stats["out_elems"] for p in parts),
                     nblocks=sum(p.stats["nblocks"] for p in parts),
                     max_chi=max(p.stats["max_chi"] for p in parts),
                     njobs=sum(p.stats["njobs"] for p in parts), n_chunks=n_chunks)
    res.options = _public_opts(opts)
    return res
