#!/usr/bin/env python
"""Benchmark of the Slater -> MPS hot path (BASELINE.json metric: sites/sec at L=1024, chi=1024).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU)
  python bench.py --impl reference ...                     the reference's algorithm on the host cores
                                                           (oracle port; TeNPy is not installed)
One "step" = one complete conversion C (resident in HBM) -> Schmidt data of all L+1 bonds and all L
block-sparse site tensors (resident in HBM).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "Slater->MPS sites/sec at L=1024, chi=1024"


def hopping_chain(L):
    H = np.zeros((L, L))
    i = np.arange(L - 1)
    H[i, i + 1] = H[i + 1, i] = -1.0
    return H


def ground_state_C(L):
    w, v = np.linalg.eigh(hopping_chain(L))
    phi = v[:, w < 0]
    return phi @ phi.T, phi.shape[1]


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during the timed region: NVML in-process (the library behind
    nvidia-smi; a query costs microseconds) or, when pynvml is missing, the nvidia-smi CLI (whose process
    start-up takes ~100 ms and perturbs the host stages of a 25 ms step)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        act = lambda bit: "Active" if (r & bit) else "Not Active"
        return [str(sm), str(mx), act(0x8), act(0x40), act(0x20), act(0x4)]   # hw, hw_thermal, sw_thermal, sw_power_cap

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                    f = [x.strip() for x in out.stdout.strip().split(",")]
                    if len(f) >= 6:
                        self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.05 if self.nvml is not None else 0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = []
        for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
            if any(s[col].lower().startswith("active") for s in self.samples):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------
def cpu_sample(L, tp, n_sites=16):
    """Times the reference algorithm for `n_sites` evenly spaced sites of the L-site chain: for each
    sampled site the marginal work of one iteration of the loops slater.py:1301-1346 (one new bond:
    eigh + enumeration; tensor data: overlap + Schur; all charge blocks: batched det)."""
    import slater_oracle as so
    C, _ = ground_state_C(L)
    trunc = so.Trunc.make(tp)
    oc = L // 2
    stride = max(1, L // n_sites)
    sites = list(range(stride // 2, L, stride))[:n_sites]
    t_total = 0.0
    for i in sites:
        if i >= oc:
            prev = so.bond_vectors_from_C(C, i, trunc, "LR" if i == oc else "R")      # not timed (carried over)
            t0 = time.perf_counter()
            new = so.bond_vectors_from_C(C, i + 1, trunc, "R")
            so.dense_tensor(so.tensor_data(new, prev, "right"))
        else:
            prev = so.bond_vectors_from_C(C, i + 1, trunc, "LR" if i + 1 == oc else "L")
            t0 = time.perf_counter()
            new = so.bond_vectors_from_C(C, i, trunc, "L")
            so.dense_tensor(so.tensor_data(new, prev, "left"))
        t_total += time.perf_counter() - t0
    return len(sites) / t_total, len(sites), t_total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tp = {"chi_max": args.chi, "svd_min": args.svd_min}
    cores = os.cpu_count()
    for _ in range(max(args.warmup, 0)):
        cpu_sample(args.L, tp, n_sites=2)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, ns, tt = cpu_sample(args.L, tp, n_sites=args.cpu_sites)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = (f"{args.cpu_sites} evenly spaced sites of the L={args.L} chain per step (one loop iteration of "
              "slater.py:1301-1346 each), oracle port of the reference (TeNPy packing excluded), NumPy/OpenBLAS")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"1D tight-binding chain L={args.L}, half filling, chi_max={args.chi}, "
                                   f"svd_min={args.svd_min:g} (BASELINE configs[4]); finite MPS; host arrays in, "
                                   "host arrays out (reference's NumPy path)"},
            "cpu_baseline": {"value": value, "unit": "sites/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    from temfpy_b200 import _lib, dist as tdist, engine, slater
    from temfpy_b200.schmidt_utils import to_stopping_condition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout when the first communicator comes up; the contract is ONE
        # JSON line on stdout, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        if not args.threads:      # the ranks share the host cores
            try:
                ncpu = len(os.sched_getaffinity(0))
            except AttributeError:
                ncpu = os.cpu_count() or 8
            args.threads = max(2, ncpu // world)
    be = engine.TorchBackend(f"cuda:{local}")
    slater._backend = be
    lib = be.lib
    L = args.L
    tp_dict = {"chi_max": args.chi, "svd_min": args.svd_min}
    tp = to_stopping_condition(tp_dict)
    C_host, N = ground_state_C(L)
    C_dev = be.from_host(C_host.ravel()) if rank == 0 else be.empty(L * L, np.float64)
    lo, hi = tdist.partition(L, world, args.chi)[rank]

    state = {}

    gather = tdist.StreamingGather(be.device) if world > 1 and not os.environ.get("TMF_SIMPLE_GATHER") else None

    def step(collect_stats=False, n_chunks=None):
        if world > 1:
            if gather is not None:
                gather.begin()
            tdist.broadcast_C(C_dev)
        res = engine.run_chain(be, C_dev, L, L, tp, N, site_lo=lo, site_hi=hi, r_sketch=args.r_sketch,
                               n_threads=args.threads, n_chunks=n_chunks if n_chunks else (args.chunks or None), lazy=True)
        if world > 1:
            if gather is not None:
                got = gather.finish(res.out_buffers())
                state["gathered"] = None if got is None else int(sum(b.numel() for b in got.values()))
            else:
                full, offs = tdist.gather_tensors(res.out_buffers())
                state["gathered"] = None if full is None else int(full.numel())
        if collect_stats:
            state["flops"] = [float(x) for x in res.flops()]
            state["out_elems"] = res.out_elems
            state["max_chi"] = max(c.max_chi for c in res.chains)
        be.sync()
        res.close()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    step(collect_stats=True)
    # ---- timed region ----------------------------------------------------------------------------
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:          # one sampler per job: eight nvidia-smi processes every 100 ms compete with the host stages
        sampler.start()
    lib.tmf_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = int(lib.tmf_launch_count(0))
    sampler.stop_flag = True
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=be.device)
    flops = torch.tensor(state["flops"], dtype=torch.float64, device=be.device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(flops)
    ms_total = float(ms.item())
    value = L * args.steps / (ms_total * 1e-3)

    # ---- per-kernel profile pass (events around every launch; separate from the timed region) ---
    # host-side stage times of one un-pipelined conversion (single chunk, explicit syncs)
    ht = []
    for _ in range(2):
        t0 = time.perf_counter()
        chain = engine.SlaterChain(be, L, tp, N, site_lo=lo, site_hi=hi, r_sketch=args.r_sketch,
                                   n_threads=args.threads)
        chain.run_modes(C_dev, L)
        t1 = time.perf_counter()
        chain.run_enumerate()
        t2 = time.perf_counter()
        chain.run_tensors(C_dev, L)
        t3 = time.perf_counter()
        be.sync()
        t4 = time.perf_counter()
        chain.close()
        ht.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    host_ms = {k: round(1e3 * float(np.mean([h[i] for h in ht])), 3) for i, k in
               enumerate(("modes_launch_to_spectra", "enumerate_and_plan_host", "tensors_enqueue", "tensors_drain"))}
    lib.tmf_prof_enable(1)
    for _ in range(2):
        step(n_chunks=1)
    import ctypes as C
    buf = C.create_string_buffer(1 << 16)
    lib.tmf_prof_report(buf, len(buf))
    lib.tmf_prof_enable(0)
    prof = {}
    for ln in buf.value.decode().strip().splitlines():
        tag, tms, cnt = ln.split()
        prof[tag] = (float(tms) / 2, int(cnt) // 2)          # per step

    # ---- FP64 peak of this GPU (MEASURED_PEAKS.json carries no FP64 figure) ------------------------
    sink = be.empty(8, np.float64)
    pms, pfl = C.c_float(0), C.c_double(0)
    lib.tmf_fp64_peak_probe(200000, be.ptr(sink), C.byref(pms), C.byref(pfl), be.stream)
    dfma_tflops = pfl.value / (pms.value * 1e-3) / 1e12
    a = torch.randn(4096, 4096, dtype=torch.float64, device=be.device)
    torch.matmul(a, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        torch.matmul(a, a)
    e1.record()
    torch.cuda.synchronize()
    dgemm_tflops = 3 * 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a
    peak = max(dfma_tflops, dgemm_tflops)

    # ---- end-to-end through the public API with host buffers ------------------------------------
    e2e = None
    if world == 1:
        slater.C_to_MPS(C_host, tp_dict, as_tenpy=False)       # warm-up (pinned allocations)
        be.h2d_bytes = be.d2h_bytes = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            mps = slater.C_to_MPS(C_host, tp_dict, as_tenpy=False)
            del mps
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        e2e = {"value": L / dt, "unit": "sites/s", "h2d_bytes_per_step": be.h2d_bytes // n_e2e,
               "d2h_bytes_per_step": be.d2h_bytes // n_e2e, "ms_per_step": dt * 1e3}
    else:
        # every rank converts its shard from host C and brings its tensors back to its host
        import torch.distributed as dist
        res = engine.run_chain(be, be.from_host(C_host.ravel()), L, L, tp, N, site_lo=lo, site_hi=hi,
                               r_sketch=args.r_sketch, n_threads=args.threads)      # warm-up (pinned allocations)
        del res
        be.h2d_bytes = be.d2h_bytes = 0
        n_e2e = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            Cd = be.from_host(C_host.ravel())
            res = engine.run_chain(be, Cd, L, L, tp, N, site_lo=lo, site_hi=hi, r_sketch=args.r_sketch,
                                   n_threads=args.threads)
            del res
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=be.device)
        be.h2d_bytes //= n_e2e
        be.d2h_bytes //= n_e2e
        by = torch.tensor([be.h2d_bytes, be.d2h_bytes], dtype=torch.float64, device=be.device)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(by)
        e2e = {"value": L / float(dt.item()), "unit": "sites/s", "h2d_bytes_per_step": int(by[0].item()),
               "d2h_bytes_per_step": int(by[1].item()), "ms_per_step": float(dt.item()) * 1e3}

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    # ---- roofline of the dominant kernel ---------------------------------------------------------
    fl = [float(x) for x in flops.tolist()]
    alg = {"eigh": fl[0], "overlap": fl[1], "schur": fl[2], "minors": fl[3]}
    # algorithmic flops attributed to each kernel family (reference's algorithm)
    family_flops = {"minors": alg["minors"], "schur": alg["schur"],
                    "gemm": alg["overlap"], "modes": alg["eigh"]}
    modes_tags = ("small_modes", "omega", "colnorm", "panel_mgs2", "svd_select", "ritz", "pivchol")
    dom = max(prof, key=lambda k: prof[k][0]) if prof else None
    roof = None
    if dom:
        dms, dcnt = prof[dom]
        if dom == "minors":
            a_fl = alg["minors"]
        elif dom == "schur":
            a_fl = alg["schur"]
        elif dom == "gemm_site":
            a_fl = alg["overlap"]
        else:                     # any kernel of the mode extraction: the reference's eigh flops
            a_fl = alg["eigh"]
        achieved = a_fl / world / (dms * 1e-3) / 1e12
        traffic = None          # DRAM bytes per launch of this kernel from the committed ncu full-set capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
            if tj and tj["launches_per_step"] == dcnt and tj["n_gpus"] == world and L == 1024 and args.chi == 1024:
                traffic = tj["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "launches_per_step": dcnt, "ms_per_step": dms,
                "algorithmic_flops_per_step": a_fl,
                "note": "FP64 pipe roofline (vector DFMA and tensor DMMA peak coincide on B200); achieved counts the "
                        "reference algorithm's flops for this stage (SURVEY 8d), the kernel itself needs far fewer",
                "peak_source": f"measured in this run: FP64 DFMA probe {dfma_tflops:.1f} TF/s, cuBLAS DGEMM 4096^3 "
                               f"{dgemm_tflops:.1f} TF/s (MEASURED_PEAKS.json has no FP64 figure)"}
    total_alg = sum(alg.values())
    whole = {"algorithmic_flops_per_step": total_alg, "achieved_tflops": total_alg / (ms_total / args.steps * 1e-3) / 1e12,
             "frac_of_fp64_peak": total_alg / (ms_total / args.steps * 1e-3) / 1e12 / (peak * world),
             "kernel_ms_per_step": {k: round(v[0], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
             "host_ms_per_step": host_ms, "minors_per_step": fl[4], "out_bytes_per_step": 8 * state["out_elems"], "max_chi": state["max_chi"]}

    cpu = None
    if world == 1 and not args.no_cpu:
        v, ns, tt = cpu_sample(L, tp_dict, n_sites=args.cpu_sites)
        cpu = {"value": v, "unit": "sites/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{ns} evenly spaced sites of the L={L} chain ({tt:.1f} s): one loop iteration of "
                         "slater.py:1301-1346 each (eigh + enumeration + overlap/Schur + batched det), oracle port, "
                         "NumPy/OpenBLAS threads = all cores"}
    line = {"metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"1D tight-binding chain L={L}, half filling, chi_max={args.chi}, "
                                   f"svd_min={args.svd_min:g} (BASELINE configs[4]); finite MPS; C resident in HBM -> "
                                   "all site tensors + Schmidt data resident in HBM",
                       "l2": f"working set {(8 * state['out_elems'] + 6e8) / 1e9:.1f} GB per step >> 126 MB L2",
                       "parallelism": (f"sites sharded over {world} GPU(s), broadcast(C) + gather(tensors) over NCCL"
                                       if world > 1 else "1 GPU") + f"; {args.chunks or 'auto (6 at >= 512 sites per GPU)'} pipeline chunks per GPU"},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roof,
            "whole_step": whole, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--L", type=int, default=1024)
    ap.add_argument("--chi", type=int, default=1024)
    ap.add_argument("--svd-min", type=float, default=1e-7)
    ap.add_argument("--r-sketch", type=int, default=48)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--chunks", type=int, default=0, help="pipeline chunks per GPU (streams + host threads)")
    ap.add_argument("--cpu-sites", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
