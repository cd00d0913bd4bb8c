"""Multi-GPU sharding of one chain (one process per GPU, ``torch.distributed`` / NCCL).

The path shards without any data-path exchange between the compute stages: every bond's Schmidt
data is a function of C alone and every site tensor needs its two adjacent bonds (reference
slater.py:1303-1309, 1328-1334), so rank r converts a contiguous, cost-balanced range of sites and
recomputes the one shared boundary bond with the same deterministic kernels (bit-identical results,
see tests).  Collectives: one broadcast of C in, one gather of the block-sparse tensors out.
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
# public multi-GPU entry point
# ---------------------------------------------------------------------------------------------
def _t(buf):
    """torch view of a backend buffer (CUDA tensor of the PyTorch backend, NumPy array of the simulator)."""
    import torch
    return buf if torch.is_tensor(buf) else torch.from_numpy(buf)


def C_to_MPS(C, trunc_par, *, ortho_center=None, spinful=None, unit_cell_width=None, dst=0, backend=None,
             n_threads=0):
    """``slater.C_to_MPS`` over all GPUs of the process group (reference slater.py:1216-1353, one process per
    GPU, ``torch.distributed`` initialised by the caller).  Every rank calls it; ``C`` is read on rank ``dst``
    only (the others may pass ``None``).  The correlation matrix is broadcast, every rank converts a contiguous,
    cost-balanced range of sites (no data-path exchange: a bond is a function of C alone), the block-sparse
    tensors are gathered on ``dst`` over NCCL / NVLink and the Schmidt tables with them.  Returns the complete
    ``BlockMPS`` on ``dst`` and ``None`` elsewhere.

    Sketch-width / fallback decisions (``engine.run_chain``) are taken for all ranks together, so that the
    boundary bond two ranks share comes out of identical kernels on both."""
    import torch
    import torch.distributed as dist
    from . import engine, slater
    from .schmidt_utils import to_stopping_condition
    world, rank = dist.get_world_size(), dist.get_rank()
    be = backend or slater._be()
    tp = to_stopping_condition(trunc_par)
    meta = [None]
    if rank == dst:
        Cp = slater._prepare_C(np.asarray(C), spinful)
        L = len(Cp)
        if unit_cell_width is None:
            unit_cell_width = L
        elif L % unit_cell_width != 0:
            raise ValueError(f"{unit_cell_width = } does not divide system size {L}")
        meta = [(L, int(np.round(np.trace(Cp))), unit_cell_width)]
    dist.broadcast_object_list(meta, src=dst)
    L, n_fermion, unit_cell_width = meta[0]
    C_dev = be.from_host(Cp.ravel()) if rank == dst else be.empty(L * L, np.float64)
    broadcast_C(_t(C_dev), src=dst)
    if rank == dst:
        slater._check_projector(Cp, be=be, Cd=C_dev)
    lo, hi = partition(L, world, tp.chi_max, ortho_center)[rank]
    opts = dict(r_sketch=48, snap=False, nested=None, device_plan=None)
    codes = {"sketch": 1, "singular": 2, "nested": 3}
    dev = _t(C_dev).device
    while True:
        res, code, err = None, 0, None
        try:
            res = engine._run_chain_once(be, C_dev, L, L, tp, n_fermion, ortho_center, lo, hi, n_threads, True, None,
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
        elif opts["nested"] is False:
            raise err or ValueError("site stage failed on another rank")
        else:
            opts["nested"] = False
    # ---- tensors -> dst (device, NVLink), Schmidt / plan tables with them ---------------------------
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
    out = engine.ChainResult(L=L, ortho_center=ortho_center or L // 2, site_lo=0, site_hi=L)
    out._pinned = pinned
    o = 0
    nblocks = max_chi = njobs = 0
    for r in range(world):
        assert o == int(offs[r])
        for st in gathered[r]:
            tab = engine.ShardTables.from_state(st.state, host[o: o + st.out_elems])
            tab.normalized()
            o += st.out_elems
            out.tables.append(tab)
            out.bonds.add([x for x in range(tab.first_bond, tab.first_bond + tab.n_bonds)
                           if st.site_lo <= x <= st.site_hi or x == out.ortho_center], tab.bond)
            out.sites.add(range(st.site_lo, st.site_hi), tab.site)
            nblocks += st.stats[0]; max_chi = max(max_chi, st.stats[1]); njobs += st.stats[2]
    out.stats = dict(out_elems=o, nblocks=nblocks, max_chi=max_chi, njobs=njobs, n_ranks=world)
    out.options = dict(opts)
    return slater._chain_to_mps(out, unit_cell_width)


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
