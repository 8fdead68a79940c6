#!/bin/bash
# bench.py (resident arm) for a list of "--kernel X [--opt ID=VAL ...]" variants given as quoted strings
set -u
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  timeout 300 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e ${WL:-} $v > gpurun_out/kv_$i.json 2> gpurun_out/kv_$i.err
  python - "$v" $i <<'PY'
import json, sys
v, i = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/kv_{i}.json"))
    print(f"{v:44s} {d['ms_per_step']:8.3f} ms/frame  kernel {d['roofline']['kernel_ms_per_frame']:8.3f} ms  launches {d['gpu_launches']}")
except Exception as e:
    print(v, "FAILED", e, open(f"gpurun_out/kv_{i}.err").read()[-600:])
PY
done
