#!/bin/bash
# runs bench.py at the given GPU counts (torchrun, one rank per GPU) and prints the headline numbers
for n in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 20 --warmup 5 $BENCH_EXTRA > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
    print("N", n, "ms/step %.2f" % d["ms_per_step"], "sites/s %.0f" % d["value"], "e2e ms %.2f" % d["e2e"]["ms_per_step"],
          "kernels", {k: v for k, v in list(d["whole_step"]["kernel_ms_per_step"].items())[:6]}, "host", d["whole_step"]["host_ms_per_step"], flush=True)
except Exception as e:
    print("N", n, "failed", e, open(f"gpurun_out/bench_n{n}.err").read()[-1500:], flush=True)
PY
done
