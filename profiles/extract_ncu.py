"""Summarise an ``ncu --set full`` report (or a launch-list CSV) into a small text file for profiles/.

    python profiles/extract_ncu.py gpurun_out/prof_minors.ncu-rep  > profiles/r01_minors_full.txt
    python profiles/extract_ncu.py --launches gpurun_out/launches.csv > profiles/r01_launches.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
    "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
    "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
]


def full(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full --clock-control none, {len(rows) - 2} launch(es)")
    for r in rows[2:]:
        print(f"\n== {r[hdr.index('Kernel Name')]}  (launch id {r[hdr.index('ID')]})")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:82s} {r[i]:>18s} {units[i]}")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
              "second": 1e3}[r[iu]]
        name = r[ik].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    print(f"# {path}: ncu --metrics gpu__time_duration.sum --clock-control none; {sum(cnt.values())} launches, "
          f"{s:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':48s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"{k:48s} {cnt[k]:8d} {tot[k]:10.3f} {100 * tot[k] / s:6.1f}%")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[1])
