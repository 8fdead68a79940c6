#!/bin/bash
# A/B of the default kernel's scheduling knobs on C2 (device-resident, no e2e / cpu baseline / parity).
set -u
mkdir -p gpurun_out
TAG=${1:-r02ab}
run() {  # name, extra args
  local name=$1; shift
  timeout 300 python bench.py --steps ${STEPS:-8} --warmup 3 --no-e2e --no-cpu-baseline --no-parity "$@" \
    > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  echo "$name: $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_${name}.json'));print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_frame'],3))" 2>&1 | tail -1)"
}
run bricks --kernel bricks
for B in 1024 768 512; do
  for G in 1 4 8 16 32; do
    run b${B}_g${G} --opt 10=$B --opt 11=$G
  done
done
