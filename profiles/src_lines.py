"""Per-source-line instruction / stall-sample shares of an ncu report captured with --import-source on.

    python profiles/src_lines.py gpurun_out/r02_minors.ncu-rep [top]
"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
hdr = None; out = []
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit():
        d = dict(zip(hdr, r))
        out.append((num(d["Instructions Executed"]), num(d["# Samples"]), int(r[0]), r[1].strip()[:100], d["Avg. Threads Executed"]))
tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print("total warp inst", tot, "samples", ts)
for o in sorted(out, key=lambda x: -x[1])[:top]:
    print(f"{o[2]:4d} inst {100*o[0]/tot:5.1f}%  smp {100*o[1]/ts:5.1f}%  thr {o[4]:>5}  {o[3]}")
