#!/bin/bash
# A/B of the default kernel's block layouts / shared-memory map on C2 + ncu captures of two of them.
set -u
mkdir -p gpurun_out
TAG=${1:-r02ab3}
cp raymarchcl_b200/libraymarch_b200.so gpurun_out/${TAG}_lib.so   # the binary the captures belong to (source-line join)
run() {
  local name=$1; shift
  timeout 300 python bench.py --steps ${STEPS:-8} --warmup 3 --no-e2e --no-cpu-baseline --no-parity "$@" \
    > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  echo "$name: $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_${name}.json'));print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_frame'],3))" 2>&1 | tail -1)"
}
if [ "${SKIP_AB:-0}" != "1" ]; then
run bricks --kernel bricks
run default
run default_again
fi
ncu_cap() {  # block, smem, group
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_persist -s 1 -c 1 -f \
    -o gpurun_out/${TAG}_persist_b$1_s$2_g$3_c2 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity \
    --opt 10=$1 --opt 12=$2 --opt 11=$3 > gpurun_out/${TAG}_ncu_b$1_g$3.log 2>&1
  echo "ncu b$1 s$2 g$3 rc=$?"
}
if [ "${SKIP_NCU:-0}" != "1" ]; then
  for v in ${NCU_CAPS:-256,0,1}; do IFS=, read b s g <<< "$v"; ncu_cap $b $s $g; done
fi
