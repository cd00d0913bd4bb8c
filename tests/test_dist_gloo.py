"""N > 1 path on CPU: two processes, gloo backend, the kernels executed by the CPU simulator.
Checks the plumbing of temfpy_b200.dist (partition, broadcast of C, gather of the block-sparse tensors)
and that the sharded conversion equals the unsharded one bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, L, chi, out_path):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import slater_oracle as so
        from temfpy_b200 import dist as tdist, engine
        from temfpy_b200.schmidt_utils import to_stopping_condition
        from tests import helpers
        from tests.hostsim import NumpyBackend
        be = NumpyBackend()
        tp = to_stopping_condition({"chi_max": chi})
        Ct = torch.zeros(L * L, dtype=torch.float64)
        n = torch.zeros(1, dtype=torch.int64)
        if rank == 0:
            Cm, nf = so.correlation_matrix(helpers.random_hamiltonian(L, 12))
            Ct.copy_(torch.from_numpy(Cm.ravel()))
            n[0] = nf
        tdist.broadcast_C(Ct)                       # C to every rank
        dist.broadcast(n, src=0)
        lo, hi = tdist.partition(L, world, chi)[rank]
        res = engine.run_chain(be, Ct.numpy(), L, L, tp, int(n[0]), site_lo=lo, site_hi=hi, n_chunks=1, lazy=True)
        buf, elems = res.out_buffers()[0]
        full, offs = tdist.gather_tensors([(torch.from_numpy(buf), elems)])       # tensors to rank 0
        lam = [res.bond(x).schmidt_values for x in range(lo, hi + 1)]
        if rank == 0:
            ref = engine.run_chain(be, Ct.numpy(), L, L, tp, int(n[0]), n_chunks=1, lazy=True)
            rbuf, relems = ref.out_buffers()[0]
            ok = int(offs[-1]) == relems and np.array_equal(full.numpy(), rbuf[:relems])
            ok = ok and all(np.array_equal(a, ref.bond(x).schmidt_values) for x, a in zip(range(lo, hi + 1), lam))
            np.save(out_path, np.array([int(ok), int(offs[-1]), relems]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("L,chi", [(24, 16)])
def test_two_rank_gloo_sharding(tmp_path, L, chi):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, _free_port(), L, chi, out), nprocs=2, join=True)
    ok, gathered, ref = np.load(out)
    assert ok == 1 and gathered == ref > 0


def _worker_public(rank, world, port, L, chi, out_path, cplx=False):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import slater_oracle as so
        from temfpy_b200 import dist as tdist, slater
        from tests import helpers
        from tests.hostsim import NumpyBackend
        be = NumpyBackend()
        Cm = None
        if rank == 0:
            Cm, _ = so.correlation_matrix(helpers.random_hamiltonian(L, 5, cplx=cplx))
        mps = tdist.C_to_MPS(Cm, {"chi_max": chi}, backend=be)          # public multi-rank entry point
        mps_nccl = tdist.C_to_MPS(Cm, {"chi_max": chi}, backend=be, host_exchange=False)
        mps2 = tdist.C_to_MPS(Cm, {"chi_max": chi}, backend=be)         # (second call: segments leased / reused)
        if rank == 0:
            assert mps.meta["stats"]["transport"] == "host segments"
            for other in (mps_nccl, mps2):
                for x in range(L + 1):
                    assert np.array_equal(mps.lams[x], other.lams[x])
                for i in range(L):
                    assert np.array_equal(mps.tensors[i].dense(), other.tensors[i].dense())
            del mps2, other
            ref = slater.C_to_MPS(Cm, {"chi_max": chi}, as_tenpy=False, _backend=be)
            ok = mps is not None and mps.L == ref.L and mps.form == ref.form
            for x in range(L + 1):
                ok = ok and np.array_equal(mps.lams[x], ref.lams[x]) and np.array_equal(mps.charges[x], ref.charges[x])
            for i in range(L):
                ok = ok and np.array_equal(mps.tensors[i].dense(), ref.tensors[i].dense())
            np.save(out_path, np.array([int(ok)]))
        else:
            assert mps is None
    finally:
        dist.destroy_process_group()


def test_two_rank_public_C_to_MPS(tmp_path):
    """dist.C_to_MPS (broadcast of C, sharded conversion, gather of tensors + tables, assembly on rank 0)
    returns the same BlockMPS, bit for bit, as the single-process slater.C_to_MPS."""
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker_public, args=(2, _free_port(), 26, 16, out), nprocs=2, join=True)
    assert np.load(out)[0] == 1


def test_two_rank_public_C_to_MPS_complex(tmp_path):
    """the same for a complex Slater determinant (real embedding on every rank, complex tensors through the
    shared host segments)."""
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker_public, args=(2, _free_port(), 14, 16, out, True), nprocs=2, join=True)
    assert np.load(out)[0] == 1


def _worker_fused(rank, world, port, L, chi, out_path):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import slater_oracle as so
        from temfpy_b200 import dist as tdist, engine
        from temfpy_b200.schmidt_utils import to_stopping_condition
        from tests import helpers
        from tests.hostsim import NumpyBackend
        be = NumpyBackend()
        tp = to_stopping_condition({"chi_max": chi})
        Cm, nf = so.correlation_matrix(helpers.random_hamiltonian(L, 12))     # (same seed on every rank)
        Ct = np.ascontiguousarray(Cm).ravel()
        fused = tdist.FusedGather(be)
        # the board: an all-gather of integers through shared memory
        got = fused.board.all_gather([10 + rank, rank * rank])
        assert got[:, 0].tolist() == [10 + r for r in range(world)] and got[:, 1].tolist() == [r * r for r in range(world)]
        oks = []
        for it in range(3):                 # the window is reused (and grown once: the first sizes are tiny)
            Lc = L if it else 8
            Cc = Ct if it else np.ascontiguousarray(so.correlation_matrix(helpers.random_hamiltonian(8, 3))[0]).ravel()
            nfc = nf if it else int(round(np.trace(Cc.reshape(8, 8))))
            lo, hi = tdist.partition(Lc, world, chi)[rank]
            res = engine.run_chain(be, Cc, Lc, Lc, tp, nfc, site_lo=lo, site_hi=hi, lazy=True, out_provider=fused)
            full, offs = fused.complete()
            if rank == 0:
                ref = engine.run_chain(be, Cc, Lc, Lc, tp, nfc, n_chunks=1, lazy=True)
                rbuf, relems = ref.out_buffers()[0]
                oks.append(int(offs[-1]) == relems and np.array_equal(np.asarray(full), rbuf[:relems]))
                ref.close()
            res.close()
        fused.close()
        if rank == 0:
            np.save(out_path, np.array([int(all(oks)), len(oks)]))
    finally:
        dist.destroy_process_group()


def test_two_rank_fused_gather(tmp_path):
    """FusedGather: both ranks' tensor stages write straight into their slice of a window owned by rank 0 (shared
    memory here, a CUDA IPC peer window on GPUs); the window then holds the unsharded result bit for bit."""
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker_fused, args=(2, _free_port(), 24, 16, out), nprocs=2, join=True)
    ok, n = np.load(out)
    assert ok == 1 and n == 3


def _worker_slotted(rank, world, port, L, chi, out_path):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import slater_oracle as so
        from temfpy_b200 import dist as tdist, engine
        from temfpy_b200.schmidt_utils import to_stopping_condition
        from tests import helpers
        from tests.hostsim import NumpyBackend
        be = NumpyBackend()
        tp = to_stopping_condition({"chi_max": chi})
        Cm, nf = so.correlation_matrix(helpers.random_hamiltonian(L, 12))
        Ct = np.ascontiguousarray(Cm).ravel()
        slots = tdist.SlottedGather(be, L, chi)
        lo, hi = tdist.partition(L, world, chi)[rank]
        oks = []
        for it in range(2):
            res = engine.run_chain(be, Ct, L, L, tp, nf, site_lo=lo, site_hi=hi, lazy=True, n_chunks=2,
                                   out_provider=slots)
            assert len(res.chains) == 2                    # two pipeline chunks per rank, each with its own slot
            win, table = slots.complete()
            if rank == 0:
                ref = engine.run_chain(be, Ct, L, L, tp, nf, n_chunks=1, lazy=True)
                rbuf, relems = ref.out_buffers()[0]
                got = np.concatenate([np.asarray(win)[o: o + n] for o, n in table])
                oks.append(len(table) == 2 * world and got.size == relems and np.array_equal(got, rbuf[:relems]))
                ref.close()
            res.close()
        slots.close()
        if rank == 0:
            np.save(out_path, np.array([int(all(oks)), len(oks)]))
    finally:
        dist.destroy_process_group()


def test_two_rank_slotted_gather(tmp_path):
    """SlottedGather: every pipeline chunk of every rank writes into its bound-sized slot of rank 0's window without
    any exchange before the completion point; the slots, in site order, hold the unsharded result bit for bit."""
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker_slotted, args=(2, _free_port(), 24, 16, out), nprocs=2, join=True)
    ok, n = np.load(out)
    assert ok == 1 and n == 2


def test_slotted_bound_covers_every_site():
    """The slot of a site range is at least the size of its tensors: 2 chi_L chi_R elements per site with
    chi <= min(chi_max, 2^x, 2^(L - x)) -- checked against actual conversions on the simulator."""
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import slater_oracle as so
    from temfpy_b200 import engine
    from temfpy_b200.schmidt_utils import to_stopping_condition
    from tests import helpers
    from tests.hostsim import NumpyBackend
    be = NumpyBackend()
    for L, chi, seed in ((20, 8, 1), (30, 64, 2), (16, 1000, 3)):
        Cm, nf = so.correlation_matrix(helpers.random_hamiltonian(L, seed))
        dmax = [min(chi, 2 ** min(x, L - x, 40)) for x in range(L + 1)]
        off = np.concatenate(([0], np.cumsum([2 * dmax[i] * dmax[i + 1] for i in range(L)])))
        for lo, hi in ((0, L), (3, 11), (L // 2, L)):
            res = engine.run_chain(be, np.ascontiguousarray(Cm).ravel(), L, L, to_stopping_condition({"chi_max": chi}), nf,
                                   site_lo=lo, site_hi=hi, n_chunks=1, lazy=True)
            assert res.out_elems <= off[hi] - off[lo], (L, chi, lo, hi)
            res.close()
