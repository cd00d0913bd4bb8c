import csv, re, sys, collections
dis, sasscsv, srcfile, kern = sys.argv[1:5]
# sequence of (line) per instruction for the kernel function
lines = open(dis).read().splitlines()
seq = []; cur = None; infun = False
for l in lines:
    if l.startswith(".text."): infun = kern in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if infun and re.match(r'\s+/\*[0-9a-f]{4}\*/', l): seq.append(cur)
rows = list(csv.reader(open(sasscsv)))
hdr = rows[1]; data = rows[2:]
iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); iSm = hdr.index("# Samples")
print("sass rows", len(data), "disasm instrs", len(seq))
agg = collections.defaultdict(lambda: [0,0,0])
for r, ln in zip(data, seq):
    a = agg[ln]; a[0] += int(r[iI]); a[1] += int(r[iT]); a[2] += int(r[iSm])
tot = sum(a[0] for a in agg.values()); smp = sum(a[2] for a in agg.values())
src = {}
for (f, n) in agg:
    pass
text = {}
import os
for root in ["/root/repo/temfpy_b200/csrc"]:
    for fn in os.listdir(root):
        if fn.endswith((".cu",".cuh",".hpp")): text[fn] = open(os.path.join(root, fn)).read().splitlines()
print("total warp inst", tot, "samples", smp)
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv)>5 else 40]:
    f, n = ln if ln else ("?", 0)
    t = text.get(f, [""]*(n+1))[n-1].strip()[:90] if n else ""
    print(f"{100*a[0]/tot:5.1f}% inst  {100*a[2]/max(smp,1):5.1f}% smp  lanes {a[1]/max(a[0],1):4.1f} | {f}:{n} {t}")
