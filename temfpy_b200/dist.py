"""Multi-GPU sharding of one chain (one process per GPU, ``torch.distributed`` / NCCL).

The path shards without any data-path exchange between the compute stages: every bond's Schmidt
data is a function of C alone and every site tensor needs its two adjacent bonds (reference
slater.py:1303-1309, 1328-1334), so rank r converts a contiguous, cost-balanced range of sites and
recomputes the one shared boundary bond with the same deterministic kernels (bit-identical results,
see tests).  Collectives: one broadcast of C in; the block-sparse tensors go out either

* into the HBM of the destination rank with the gather fused into the minors kernel -- the kernel's output pointer is
  a slice of a CUDA-IPC peer window on the destination GPU (:class:`SlottedGather`: slots sized by an upper bound, no
  exchange before the completion point; :class:`FusedGather`: exact slices, sizes exchanged through shared memory
  after the enumeration stage), or by NCCL send / recv (:class:`StreamingGather`, :func:`gather_tensors`), or
* into the host memory of the destination *process* over every GPU's own PCIe link (:class:`HostExchange`: shared
  pinned segments, tables included, zero copy on the destination) -- what :func:`C_to_MPS` uses.
"""
from __future__ import annotations

import numpy as np


def site_costs(L: int, chi_max: int | None, ortho_center: int | None = None) -> np.ndarray:
    """Relative cost model per site, fitted to measured shard times on B200 (8 ranks, L = 1024): the mode
    extraction of a block of n sites costs ~ (150 + n) (sketch + skinny GEMMs + latency-bound panels, no
    O(n^3) eigen-solver any more), the tensor entries ~ chi_bra * chi_ket (chi saturates at chi_max,
    2^distance near the ends)."""
    oc = ortho_center or L // 2
    i = np.arange(L)
    n = np.where(i >= oc, L - i, i + 1).astype(float)
    dist = np.minimum(i + 1, L - i).astype(float)
    cap = float(chi_max) if chi_max else 1024.0
    chi = np.minimum(cap, 2.0 ** np.minimum(dist, 40))
    return (150.0 + n) * (0.35 + 0.65 * (chi / cap) ** 2)


def partition(L: int, world: int, chi_max: int | None = None, ortho_center: int | None = None,
              lo: int = 0, hi: int | None = None, weights=None):
    """Contiguous ranges covering [lo, hi) with (nearly) equal summed cost, or with costs proportional to
    ``weights`` (one per range)."""
    hi = L if hi is None else hi
    c = np.cumsum(site_costs(L, chi_max, ortho_center)[lo:hi])
    n = hi - lo
    world = max(1, min(world, n))
    if weights is None or len(weights) != world:
        targets = np.arange(1, world) / world
    else:
        w = np.asarray(weights, dtype=float)
        targets = np.cumsum(w)[:-1] / w.sum()
    cuts = [0]
    for t in targets:
        cuts.append(int(np.searchsorted(c, c[-1] * t)))
    cuts.append(n)
    for r in range(1, world + 1):           # every rank gets at least one site
        cuts[r] = max(cuts[r], cuts[r - 1] + 1)
    cuts[-1] = n
    for r in range(world - 1, 0, -1):
        cuts[r] = min(cuts[r], cuts[r + 1] - 1)
    return [(lo + cuts[r], lo + cuts[r + 1]) for r in range(world)]


def broadcast_C(C_dev, src=0):
    """Correlation matrix to every rank over NCCL (8 MiB at L = 1024)."""
    import torch.distributed as dist
    dist.broadcast(C_dev, src=src)


MAX_PARTS = 16


def gather_tensors(parts, dst=0):
    """Gathers the ranks' block-sparse tensor buffers on ``dst`` (variable sizes -> grouped
    point-to-point sends over NVLink).  ``parts``: list of (device buffer, n_elements) in site order (the
    pipeline chunks of this rank; sent one by one, no concatenation copy).  Returns (buffer, offsets per
    rank) on dst, (None, None) elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if not isinstance(parts, (list, tuple)):
        raise TypeError("parts must be a list of (buffer, n_elements)")
    if len(parts) > MAX_PARTS:
        raise ValueError("too many pipeline chunks for one gather")
    dev = parts[0][0].device
    sizes_h = np.zeros((world, MAX_PARTS), dtype=np.int64)
    for i, (_, n) in enumerate(parts):
        sizes_h[rank, i] = int(n)
    sizes = torch.from_numpy(sizes_h).to(dev)
    dist.all_reduce(sizes)
    sizes = sizes.cpu().numpy()
    per_rank = sizes.sum(axis=1)
    offs = np.concatenate(([0], np.cumsum(per_rank))).astype(np.int64)
    if rank == dst:
        full = torch.empty(int(offs[-1]), dtype=parts[0][0].dtype, device=dev)
        ops = []
        for r in range(world):
            o = int(offs[r])
            for i in range(MAX_PARTS):
                n = int(sizes[r, i])
                if n == 0:
                    continue
                if r == dst:
                    full[o: o + n].copy_(parts[i][0][:n])
                else:
                    ops.append(dist.P2POp(dist.irecv, full[o: o + n], r))
                o += n
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return full, offs
    ops = [dist.P2POp(dist.isend, buf[:n], dst) for buf, n in parts if n]
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    return None, None


class StreamingGather:
    """Gather of the ranks' tensor buffers on rank ``dst`` that does not wait for the slowest rank before
    the first byte moves: every (dst, r) pair has its own two-rank NCCL communicator, a helper thread on
    ``dst`` receives rank r's sizes and posts the matching receives as soon as *that* rank is done, so the
    transfers of the early ranks overlap the kernels of the late ones (the all-reduce of the sizes in
    ``gather_tensors`` made every transfer start after the slowest rank; at 8 GPUs that cost ~2 ms of a
    15 ms step).  ``begin()`` at the start of a step, ``finish(parts)`` after the local conversion."""

    def __init__(self, device, dst=0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world, self.rank, self.dst, self.device = dist.get_world_size(), dist.get_rank(), dst, device
        self.groups = {}
        for r in range(self.world):          # new_group is collective: every rank creates every pair group
            if r != dst:
                self.groups[r] = dist.new_group(ranks=sorted((dst, r)))
        self.stream = torch.cuda.Stream(device=device) if self.rank == dst else None
        self.thread = None
        self.result = None
        self.error = None

    def begin(self):
        if self.rank != self.dst:
            return
        import threading
        self.result, self.error = {}, None

        def serve():
            t, d = self.torch, self.dist
            try:
                with t.cuda.stream(self.stream):
                    pending = []
                    sizes = {}
                    for r in self.groups:                      # one tiny message per rank: its chunk sizes
                        st = t.zeros(MAX_PARTS, dtype=t.int64, device=self.device)
                        w = d.irecv(st, src=r, group=self.groups[r])
                        sizes[r] = (st, w)
                    waiting = list(self.groups)
                    while waiting:                               # whichever rank reports first is served first
                        for r in list(waiting):
                            st, w = sizes[r]
                            if w.is_completed():
                                w.wait()
                                n = st.cpu().numpy()
                                buf = t.empty(int(n.sum()), dtype=t.float64, device=self.device)
                                o = 0
                                for k in n:
                                    if k:
                                        pending.append(d.irecv(buf[o: o + int(k)], src=r, group=self.groups[r]))
                                        o += int(k)
                                self.result[r] = buf
                                waiting.remove(r)
                        if waiting:
                            import time
                            time.sleep(0.0001)
                    for w in pending:
                        w.wait()
                    self.stream.synchronize()
            except Exception as e:       # surfaced by finish()
                self.error = e

        self.thread = threading.Thread(target=serve, name="tmf-gather", daemon=True)
        self.thread.start()

    def finish(self, parts):
        """parts: list of (device buffer, n_elements) of this rank.  Returns {rank: buffer} on dst, None elsewhere."""
        t, d = self.torch, self.dist
        if self.rank != self.dst:
            if len(parts) > MAX_PARTS:
                raise ValueError("too many pipeline chunks for one gather")
            n = np.zeros(MAX_PARTS, dtype=np.int64)
            for i, (_, k) in enumerate(parts):
                n[i] = int(k)
            g = self.groups[self.rank]
            ws = [d.isend(t.from_numpy(n).to(self.device), dst=self.dst, group=g)]
            ws += [d.isend(buf[:k], dst=self.dst, group=g) for buf, k in parts if k]
            for w in ws:
                w.wait()
            return None
        own = t.cat([b[:k] for b, k in parts]) if len(parts) > 1 else parts[0][0][: parts[0][1]]
        self.thread.join()
        if self.error is not None:
            raise self.error
        out = dict(self.result)
        out[self.dst] = own
        return out


# ---------------------------------------------------------------------------------------------
# gather fused into the tensor kernels: peer window in the destination rank's HBM
# ---------------------------------------------------------------------------------------------
def _shm_dir():
    import os
    import tempfile
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()


class HostBoard:
    """All-gather of a few integers between the ranks of one node through POSIX shared memory: every rank owns a
    row (sequence number + two payload banks), publishes its values and spins until all rows carry the sequence
    number.  A few microseconds, no GPU work, no NCCL launch -- used where the ranks must agree on sizes in the
    middle of a conversion (offsets inside the peer window) and as the completion point of a fused gather."""
    WIDTH = 12

    def __init__(self, dst=0, timeout=120.0):
        import os
        import torch.distributed as dist
        self.world, self.rank, self.timeout = dist.get_world_size(), dist.get_rank(), timeout
        name = [None]
        if self.rank == dst:
            name = [os.path.join(_shm_dir(), f"tmfpy_board_{os.getpid()}_{id(self):x}")]
            np.zeros((self.world, 32), dtype=np.int64).tofile(name[0])
        dist.broadcast_object_list(name, src=dst)
        self.path, self.owner = name[0], self.rank == dst
        self.arr = np.memmap(self.path, dtype=np.int64, mode="r+", shape=(self.world, 32))
        self.seq = 0
        dist.barrier()                      # everybody has mapped the file: the name can go
        if self.owner:
            os.unlink(self.path)

    def all_gather(self, vals):
        import time
        vals = [int(v) for v in vals]
        assert len(vals) <= self.WIDTH
        self.seq += 1
        bank = 2 + 14 * (self.seq & 1)
        row = self.arr[self.rank]
        row[bank: bank + len(vals)] = vals
        row[0] = self.seq                   # published after the payload (x86 keeps the store order)
        t0 = time.perf_counter()
        spins = 0
        while int(self.arr[:, 0].min()) < self.seq:
            spins += 1
            if spins & 1023 == 0:
                if time.perf_counter() - t0 > self.timeout:
                    raise RuntimeError("HostBoard.all_gather: a rank did not arrive (timeout)")
                time.sleep(0)
        return np.array(self.arr[:, bank: bank + self.WIDTH])


class _RawBuffer:
    """Device memory that PyTorch did not allocate in this process (a slice of a peer window)."""

    def __init__(self, ptr, nbytes):
        self._ptr, self.nbytes = int(ptr), int(nbytes)

    def data_ptr(self):
        return self._ptr


class PeerWindow:
    """``nbytes`` of HBM on rank ``dst`` that every rank of the node can address: allocated by ``dst``, exported
    as a CUDA IPC handle, mapped by the others (``tmf_ipc_export`` / ``tmf_ipc_open``).  With the simulator
    backend (CPU tests) the window is a shared-memory file instead."""

    def __init__(self, be, nbytes, dst=0):
        import ctypes as C
        import os
        import torch.distributed as dist
        from ._lib import check
        self.be, self.dst, self.nbytes = be, dst, int(nbytes)
        self.rank = dist.get_rank()
        self.cuda = hasattr(be, "torch")
        self.base = None
        info = [None]
        if self.cuda:
            if self.rank == dst:
                try:
                    self.buf = be.empty(self.nbytes, np.uint8)
                    h, off = C.create_string_buffer(64), C.c_int64()
                    check(be.lib, be.lib.tmf_ipc_export(be.ptr(self.buf), h, C.byref(off)))
                    info = [(h.raw, int(off.value))]
                except Exception as exc:          # the others must not wait for a handle that never comes
                    info = [("failed", repr(exc))]
            dist.broadcast_object_list(info, src=dst)
            if info[0][0] == "failed":
                raise RuntimeError(f"peer window: export failed on rank {dst}: {info[0][1]}")
            if self.rank == dst:
                self.ptr = be.ptr(self.buf)
            else:
                base = C.c_void_p()
                check(be.lib, be.lib.tmf_ipc_open(info[0][0], C.byref(base)))
                self.base = base.value
                self.ptr = self.base + info[0][1]
        else:
            if self.rank == dst:
                info = [os.path.join(_shm_dir(), f"tmfpy_win_{os.getpid()}_{id(self):x}")]
                with open(info[0], "wb") as f:
                    f.truncate(max(self.nbytes, 8))
            dist.broadcast_object_list(info, src=dst)
            self.buf = np.memmap(info[0], dtype=np.float64, mode="r+", shape=(max(self.nbytes // 8, 1),))
            dist.barrier()
            if self.rank == dst:
                os.unlink(info[0])

    def slice(self, off_bytes, nbytes):
        """Backend buffer for [off_bytes, off_bytes + nbytes) of the window."""
        assert 0 <= off_bytes and off_bytes + nbytes <= self.nbytes
        if not self.cuda:
            return self.buf[off_bytes // 8: (off_bytes + nbytes) // 8]
        if self.rank == self.dst:
            return self.buf[off_bytes: off_bytes + nbytes].view(self.be.torch.float64)
        return _RawBuffer(self.ptr + off_bytes, nbytes)

    def close(self):
        if self.cuda and self.base is not None:
            self.be.lib.tmf_ipc_close(self.base)
            self.base = None
        self.buf = None


class FusedGather:
    """Gather of a sharded conversion without a gather step: ``engine.run_chain(..., out_provider=FusedGather)``
    makes every rank's tensor kernels store their output directly into the rank's slice of a peer window on
    ``dst`` (P2P stores over NVLink, overlapped with the minors arithmetic tile by tile).  The slices follow each
    other in rank = site order, so ``dst`` ends up with the same contiguous buffer ``gather_tensors`` returns.

    The slice offsets need every rank's tensor size, which is known after its enumeration stage: the sizes are
    exchanged through a :class:`HostBoard` (no device work).  A rank that leaves the conversion early (retry with
    other options, ``engine._Retry``) publishes -1 and the others follow it out."""

    def __init__(self, be, dst=0):
        import torch.distributed as dist
        self.be, self.dst = be, dst
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.board = HostBoard(dst)
        self.win = None
        self.sizes = None

    def _ensure(self, nbytes):
        if self.win is not None and self.win.nbytes >= nbytes:
            return
        if self.win is not None:
            if hasattr(self.be, "sync"):
                self.be.sync()
            self.board.all_gather([0])              # nobody still writes into the old window
            self.win.close()
        cap = max(int(nbytes * 1.25), 1 << 20)
        self.win = PeerWindow(self.be, (cap + 255) & ~255, self.dst)

    def __call__(self, chain):
        from . import engine
        n = int(chain.tensor_doubles())
        sizes = self.board.all_gather([n])[:, 0]
        if (sizes < 0).any():
            raise engine._Retry("peer", ValueError("another rank restarts the conversion with other options"))
        self._ensure(8 * int(sizes.sum()))          # collective: every rank sees the same sizes
        self.sizes = sizes
        return self.win.slice(8 * int(sizes[: self.rank].sum()), 8 * n)

    def abort(self):
        self.board.all_gather([-1])

    def complete(self):
        """Completion point: call after the local stream is synchronised; returns (buffer, offsets) on ``dst`` once
        every rank's kernels have finished, (None, None) elsewhere."""
        self.board.all_gather([1])
        if self.rank != self.dst:
            return None, None
        offs = np.concatenate(([0], np.cumsum(self.sizes))).astype(np.int64)
        return self.win.slice(0, 8 * int(offs[-1])), offs

    def close(self):
        if self.win is not None:
            self.win.close()
            self.win = None


class SlottedGather:
    """:class:`FusedGather` without the exchange in the middle of the conversion: the window on ``dst`` is laid out by
    an upper bound of every site's tensor (2 chi_L chi_R elements with chi <= min(chi_max, 2^x, 2^(L-x))), so the
    slot of a pipeline chunk follows from its site range alone.  Every chunk of every rank writes at its fixed
    offset as soon as its own tensor stage starts -- several pipeline chunks per rank keep ``dst``'s NVLink port busy
    while later chunks are still in their mode stage -- and the actual sizes are published once, at the completion
    point.  ``dst`` ends up with all tensors in its HBM plus the table of (offset, size) per chunk in site order.
    The bound costs address space, not traffic (17 GB for L = chi_max = 1024); conversions without ``chi_max`` or
    beyond ``max_bytes`` use :class:`FusedGather`."""
    multi_chunk = True

    def __init__(self, be, L, chi_max, es=1, dst=0, max_bytes=48 << 30):
        import torch.distributed as dist
        if chi_max is None:
            raise ValueError("SlottedGather needs chi_max")
        self.be, self.dst, self.es = be, dst, int(es)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        dmax = [min(int(chi_max), 2 ** min(x, L - x, 40)) for x in range(L + 1)]
        per_site = np.array([2 * dmax[i] * dmax[i + 1] for i in range(L)], dtype=np.int64) * self.es
        self.off = np.concatenate(([0], np.cumsum(per_site))).astype(np.int64)      # doubles
        cap = 8 * int(self.off[-1])
        if cap > max_bytes:
            raise ValueError(f"SlottedGather: window of {cap >> 30} GiB exceeds the limit")
        self.board = HostBoard(dst)
        self.win = PeerWindow(be, (cap + 255) & ~255, dst)
        self.chunks = []

    def __call__(self, chain):
        n = int(chain.tensor_doubles())
        lo, hi = int(chain.site_lo), int(chain.site_hi)
        if n > int(self.off[hi] - self.off[lo]):
            raise RuntimeError("SlottedGather: a chunk exceeds its bound")
        self.chunks.append((lo, n))
        return self.win.slice(8 * int(self.off[lo]), 8 * n)

    def abort(self):
        pass                                    # nobody waits for this rank before the completion point

    def complete(self):
        """After the local streams are synchronised: (window, [(offset, doubles), ...] in site order) on ``dst`` once
        every rank's kernels have finished, (None, None) elsewhere."""
        mine, self.chunks = sorted(self.chunks), []
        if len(mine) > HostBoard.WIDTH // 2:
            raise ValueError("too many pipeline chunks for one gather")
        payload = [-1] * HostBoard.WIDTH
        for i, (lo, n) in enumerate(mine):
            payload[2 * i], payload[2 * i + 1] = lo, n
        got = self.board.all_gather(payload)
        if self.rank != self.dst:
            return None, None
        table = sorted((int(got[r, 2 * i]), int(got[r, 2 * i + 1])) for r in range(self.world)
                       for i in range(HostBoard.WIDTH // 2) if got[r, 2 * i] >= 0)
        return self.win.slice(0, self.win.nbytes), [(int(self.off[lo]), n) for lo, n in table]

    def close(self):
        if self.win is not None:
            self.win.close()
            self.win = None


# ---------------------------------------------------------------------------------------------
# results to the destination *process*: shared pinned host segments, one PCIe link per GPU
# ---------------------------------------------------------------------------------------------
class _Segment:
    """A file in shared memory mapped into this process.  Layout: int64[0] lease flag (1 while the destination
    process holds views of the contents), int64[1] length of the pickled directory that starts at byte 64, data
    from byte ``DATA`` on.  The owner (the rank that fills it) registers the mapping with the CUDA driver so that
    its device -> host copies run at PCIe speed and asynchronously."""
    DATA = 1 << 16

    def __init__(self, path, nbytes, lib=None, create=False):
        import mmap
        import os
        self.path, self.nbytes, self.lib, self.owner = path, int(nbytes), lib, create
        fd = os.open(path, os.O_RDWR | (os.O_CREAT if create else 0), 0o600)
        try:
            if create:
                os.ftruncate(fd, self.nbytes)
            self.mm = mmap.mmap(fd, self.nbytes)
        finally:
            os.close(fd)
        self.arr = np.frombuffer(self.mm, dtype=np.uint8)
        self.head = self.arr[:16].view(np.int64)
        self.registered = False
        if create:
            self.head[:] = 0
            if lib is not None and lib.tmf_is_cuda():
                from ._lib import check
                check(lib, lib.tmf_host_register(self.arr.ctypes.data, self.nbytes))
                self.registered = True

    def close(self):
        import os
        if self.registered:
            self.lib.tmf_host_unregister(self.arr.ctypes.data)
            self.registered = False
        if self.owner:
            try:
                os.unlink(self.path)
            except OSError:
                pass
            self.owner = False


class HostExchange:
    """Moves every rank's shard (tensors + Schmidt / plan tables) into the destination process without passing
    through the destination GPU: each rank copies device -> its own shared pinned segment over its own PCIe link
    (all links in parallel), publishes the segment through the :class:`HostBoard`, and the destination process maps
    the segments and wraps the arrays in place (zero copy).  A segment is leased to the destination until the last
    NumPy view of it is garbage-collected there; the filling rank then reuses it (registration is paid once)."""

    def __init__(self, be, dst=0):
        import atexit
        import os
        import torch.distributed as dist
        self.be, self.dst = be, dst
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.board = HostBoard(dst)
        tok = [f"tmfpy_seg_{os.getpid()}_{id(self):x}" if self.rank == dst else None]
        dist.broadcast_object_list(tok, src=dst)
        self.prefix = os.path.join(_shm_dir(), tok[0])
        self.mine = []          # segments this rank fills
        self.peers = {}         # dst: (rank, index) -> _Segment
        atexit.register(self.close)
        # every rank must be able to create, map and register shared memory; otherwise all ranks use the NCCL gather
        ok = 1
        try:
            probe = _Segment(f"{self.prefix}_probe_r{self.rank}", 1 << 16, be.lib, create=True)
            probe.close()
        except Exception:
            ok = 0
        self.usable = bool(self.board.all_gather([ok])[:, 0].min())

    def _acquire(self, need):
        for i, seg in enumerate(self.mine):
            if seg.nbytes >= need and int(seg.head[0]) == 0:
                return i, seg
        for i, seg in enumerate(self.mine):           # too small and free: replace
            if int(seg.head[0]) == 0:
                seg.close()
                self.mine[i] = None
        self.mine = [m for m in self.mine if m is not None]
        i = getattr(self, "_next", 0)
        self._next = i + 1
        cap = (int(need * 1.25) + (1 << 21)) & ~((1 << 21) - 1)
        seg = _Segment(f"{self.prefix}_r{self.rank}_{i}", cap, self.be.lib, create=True)
        seg.index = i
        self.mine.append(seg)
        return i, seg

    def send(self, res, extra=None):
        """Every rank: shard ``res`` (``engine.DeviceChainResult``) -> its segment; returns on ``dst`` the list,
        in rank order, of ``(directory, root array)`` per rank, ``None`` elsewhere."""
        import pickle
        from . import engine
        be, lib = self.be, self.be.lib
        chains = res.chains
        tabs = [engine.ShardTables(c, None, want_sites=True) for c in chains]
        states = [t.state() for t in tabs]
        need = _Segment.DATA
        for c, st in zip(chains, states):
            need += 8 * c.es * c.out_elems + 256
            for v in st.values():
                need += (len(v) if isinstance(v, (bytes, bytearray)) else getattr(v, "nbytes", 64)) + 64
        idx, seg = self._acquire(need)
        seg.head[0] = 1                                   # leased from now on
        o = _Segment.DATA
        directory = []
        cuda = hasattr(be, "torch")
        for c, st in zip(chains, states):
            nb = 8 * c.es * c.out_elems
            if cuda:
                from ._lib import check
                check(lib, lib.tmf_copy_d2h_async(seg.arr.ctypes.data + o, be.ptr(c._buffers["out"]), nb, be.stream))
                be.d2h_bytes += nb
            else:
                seg.arr[o: o + nb] = np.asarray(c._buffers["out"]).view(np.uint8)[:nb]
            ent = dict(out=(o, c.out_elems, c.es), site_lo=c.site_lo, site_hi=c.site_hi,
                       stats=(c.nblocks, c.max_chi, c.njobs), arrays={}, scalars={})
            o = (o + nb + 63) & ~63
            for k, v in st.items():                       # tables: written while the tensor copy is in flight
                if isinstance(v, (bytes, bytearray)):
                    v = np.frombuffer(v, dtype=np.uint8)
                if isinstance(v, np.ndarray):
                    v = np.ascontiguousarray(v)
                    seg.arr[o: o + v.nbytes] = v.reshape(-1).view(np.uint8)
                    ent["arrays"][k] = (o, v.dtype.str, v.shape)
                    o = (o + v.nbytes + 63) & ~63
                else:
                    ent["scalars"][k] = v
            directory.append(ent)
        blob = pickle.dumps(dict(chunks=directory, extra=extra), protocol=pickle.HIGHEST_PROTOCOL)
        assert 64 + len(blob) <= _Segment.DATA and o <= seg.nbytes
        seg.arr[64: 64 + len(blob)] = np.frombuffer(blob, dtype=np.uint8)
        seg.head[1] = len(blob)
        be.sync()                                          # the tensor copy has landed
        got = self.board.all_gather([seg.index, seg.nbytes])
        if self.rank != self.dst:
            return None
        out = []
        for r in range(self.world):
            key, nbytes = (r, int(got[r, 0])), int(got[r, 1])
            ps = self.peers.get(key)
            if ps is None or ps.nbytes != nbytes:
                ps = seg if r == self.rank else _Segment(f"{self.prefix}_r{r}_{key[1]}", nbytes)
                self.peers[key] = ps
            root = np.frombuffer(ps.mm, dtype=np.uint8)    # fresh root: every view of this result hangs off it
            import weakref
            weakref.finalize(root, _release, ps.head)
            d = pickle.loads(root[64: 64 + int(ps.head[1])].tobytes())
            out.append((d, root))
        return out

    def close(self):
        for seg in self.mine:
            seg.close()
        self.mine = []


def _release(head):
    head[0] = 0


def _tables_from_segment(ent, root):
    from . import engine
    st = dict(ent["scalars"])
    for k, (o, dt, shape) in ent["arrays"].items():
        n = int(np.prod(shape)) * np.dtype(dt).itemsize
        st[k] = root[o: o + n].view(np.dtype(dt)).reshape(shape)
    if "plans" in st:
        st["plans"] = st["plans"].tobytes()
    o, n, es = ent["out"]
    host = root[o: o + 8 * es * n].view(np.complex128 if es == 2 else np.float64)
    return engine.ShardTables.from_state(st, host)


# ---------------------------------------------------------------------------------------------
# public multi-GPU entry point
# ---------------------------------------------------------------------------------------------
def _t(buf):
    """torch view of a backend buffer (CUDA tensor of the PyTorch backend, NumPy array of the simulator)."""
    import torch
    return buf if torch.is_tensor(buf) else torch.from_numpy(buf)


def C_to_MPS(C, trunc_par, *, ortho_center=None, spinful=None, unit_cell_width=None, dst=0, backend=None,
             n_threads=0, host_exchange=True):
    """``slater.C_to_MPS`` over all GPUs of the process group (reference slater.py:1216-1353, one process per
    GPU, ``torch.distributed`` initialised by the caller).  Every rank calls it; ``C`` is read on rank ``dst``
    only (the others may pass ``None``).  The correlation matrix is broadcast, every rank converts a contiguous,
    cost-balanced range of sites (no data-path exchange: a bond is a function of C alone).  The result is wanted in
    the *host* memory of ``dst``: by default every GPU copies its shard (tensors and Schmidt tables) into a shared
    pinned host segment over its own PCIe link and ``dst`` wraps the segments in place (:class:`HostExchange`; all
    ranks on one node); ``host_exchange=False`` gathers the tensors in ``dst``'s HBM over NCCL / NVLink first.
    (Callers that want the tensors in ``dst``'s HBM use ``engine.run_chain(..., out_provider=FusedGather(...))``.)
    Returns the complete ``BlockMPS`` on ``dst`` and ``None`` elsewhere.

    Sketch-width / fallback decisions (``engine.run_chain``) are taken for all ranks together, so that the
    boundary bond two ranks share comes out of identical kernels on both."""
    import torch
    import torch.distributed as dist
    from . import engine, slater
    from .schmidt_utils import to_stopping_condition
    world, rank = dist.get_world_size(), dist.get_rank()
    be = backend or slater._be()
    tp = to_stopping_condition(trunc_par)
    if host_exchange:           # (shared pinned segments must work on every rank, else the NCCL gather)
        key = (id(be), dst)
        if key not in _exchanges:
            _exchanges[key] = HostExchange(be, dst)
        host_exchange = _exchanges[key].usable
    meta = [None]
    if rank == dst:
        Cp = slater._prepare_C(np.asarray(C), spinful)
        L = len(Cp)
        if unit_cell_width is None:
            unit_cell_width = L
        elif L % unit_cell_width != 0:
            raise ValueError(f"{unit_cell_width = } does not divide system size {L}")
        cplx = bool(np.iscomplexobj(Cp))
        meta = [(L, int(np.round(np.trace(Cp).real)), unit_cell_width, cplx)]
    dist.broadcast_object_list(meta, src=dst)
    L, n_fermion, unit_cell_width, cplx = meta[0]
    # complex Slater determinants: the kernels work on the 2L x 2L real embedding (cuts at 2x), see slater.C_to_MPS
    Lc = 2 * L if cplx else L
    if rank == dst:
        Ce = slater.embed_complex(Cp) if cplx else Cp
        C_dev = be.from_host(Ce.ravel())
    else:
        C_dev = be.empty(Lc * Lc, np.float64)
    broadcast_C(_t(C_dev), src=dst)
    if rank == dst:
        slater._check_projector(Ce, be=be, Cd=C_dev)
    lo, hi = partition(L, world, tp.chi_max, ortho_center)[rank]
    opts = dict(r_sketch=96 if cplx else 48, snap=False, nested=engine.default_nested(None, tp, cplx), device_plan=None,
                cplx=cplx)
    codes = {"sketch": 1, "singular": 2, "nested": 3, "peer": 0}
    dev = _t(C_dev).device
    while True:
        res, code, err = None, 0, None
        try:
            res = engine._run_chain_once(be, C_dev, Lc, L, tp, n_fermion, ortho_center, lo, hi, n_threads, True, None,
                                         True, opts)
        except engine._Retry as rt:
            code, err = codes[rt.kind], rt.err
        flag = torch.tensor([code], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        worst = int(flag.item())
        if worst == 0:
            break
        if res is not None:
            res.close()
        if worst == 1:
            wider = [r for r in engine.SKETCH_WIDTHS if r > opts["r_sketch"]]
            if not wider:
                raise err or ValueError("range sketch too narrow on another rank")
            opts["r_sketch"] = wider[0]
        elif worst == 2 and not opts["snap"]:
            opts["snap"] = True
        elif opts["nested"] is False or cplx:
            raise err or ValueError("site stage failed on another rank")
        else:
            opts["nested"] = False
    # ---- results -> dst ------------------------------------------------------------------------------------
    out = engine.ChainResult(L=L, ortho_center=ortho_center or L // 2, site_lo=0, site_hi=L)
    nblocks = max_chi = njobs = o = 0

    def add(tab, site_lo, site_hi, stats):
        nonlocal nblocks, max_chi, njobs
        tab.normalized()
        out.tables.append(tab)
        out.bonds.add([x for x in range(tab.first_bond, tab.first_bond + tab.n_bonds)
                       if site_lo <= x <= site_hi or x == out.ortho_center], tab.bond)
        out.sites.add(range(site_lo, site_hi), tab.site)
        nblocks += stats[0]; max_chi = max(max_chi, stats[1]); njobs += stats[2]

    if host_exchange:
        # every GPU copies its shard into a shared pinned host segment over its own PCIe link; dst maps the segments
        # and wraps tensors and tables in place
        key = (id(be), dst)
        ex = _exchanges[key]
        got = ex.send(res)
        res.close()
        if rank != dst:
            return None
        for d, root in got:
            for ent in d["chunks"]:
                add(_tables_from_segment(ent, root), ent["site_lo"], ent["site_hi"], ent["stats"])
                o += ent["out"][1]
    else:
        # tensors -> dst's HBM over NCCL / NVLink, pinned device -> host copy there; tables pickled alongside
        parts = [(_t(b), n) for b, n in res.out_buffers()]
        full, offs = gather_tensors(parts, dst=dst)
        states = [ShardState(engine.ShardTables(c, None, want_sites=True).state(), c.out_elems, c.site_lo, c.site_hi,
                             (c.nblocks, c.max_chi, c.njobs)) for c in res.chains]
        gathered = [None] * world if rank == dst else None
        dist.gather_object(states, gathered, dst=dst)
        res.close()
        if rank != dst:
            return None
        if hasattr(be, "to_host_async") and full.is_cuda:
            pinned = be.to_host_async(full)
            be.sync()
            host = pinned.numpy()
        else:
            pinned, host = None, full.numpy()
        out._pinned = pinned
        es = 2 if cplx else 1
        for r in range(world):
            assert es * o == int(offs[r])
            for st in gathered[r]:
                blk = host[es * o: es * (o + st.out_elems)]
                add(engine.ShardTables.from_state(st.state, blk.view(np.complex128) if cplx else blk),
                    st.site_lo, st.site_hi, st.stats)
                o += st.out_elems
    out.stats = dict(out_elems=o, nblocks=nblocks, max_chi=max_chi, njobs=njobs, n_ranks=world,
                     transport="host segments" if host_exchange else "nccl gather")
    out.options = engine._public_opts(opts)
    return slater._chain_to_mps(out, unit_cell_width)


_exchanges = {}


class ShardState:
    """What a rank sends to ``dst`` for each of its pipeline chunks (picklable)."""

    def __init__(self, state, out_elems, site_lo, site_hi, stats):
        self.state, self.out_elems, self.site_lo, self.site_hi, self.stats = state, out_elems, site_lo, site_hi, stats


def H_to_MPS(H, trunc_par, **kw):
    """``slater.H_to_MPS`` over all GPUs of the process group (reference slater.py:1568-1627): the one-off
    ``eigh(H)`` + ``C = Phi Phi^T`` run on rank ``dst``, the conversion on all ranks (see ``C_to_MPS``)."""
    import torch.distributed as dist
    from . import slater
    C = None
    if dist.get_rank() == kw.get("dst", 0):
        C, _ = slater.correlation_matrix(H, _backend=kw.get("backend"))
    return C_to_MPS(C, trunc_par, **kw)
