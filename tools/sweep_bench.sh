#!/bin/bash
# usage: tools/sweep_bench.sh "<env assignments>" chunks...   -> ms/step of the device-resident metric and of e2e
envs="$1"; shift
for nc in "$@"; do
  env $envs python bench.py --steps 20 --warmup 5 --no-cpu --chunks $nc 2>/dev/null > /tmp/sweep.json
  python - "$envs" "$nc" <<'PY'
import json, sys
d = json.load(open("/tmp/sweep.json"))
print("env", sys.argv[1] or "-", "chunks", sys.argv[2], "ms/step %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], flush=True)
PY
done
