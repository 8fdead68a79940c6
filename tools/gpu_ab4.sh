#!/bin/bash
# A/B of library variants x option sets: tools/gpu_ab4.sh "name[:ID=VALUE,...]" ...
set -u
mkdir -p gpurun_out
for spec in "$@"; do
  v=${spec%%:*}; o=""; tag=$v
  if [[ "$spec" == *:* ]]; then for kv in $(echo "${spec#*:}" | tr ',' ' '); do o="$o --opt $kv"; done; tag=$(echo "$spec" | tr ':=,' '___'); fi
  RAYMARCH_B200_LIB=$PWD/build_ab/$v.so timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} $o \
    > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{v}.json"))
    print(f"{v:28s} {d['ms_per_step']:8.3f} ms/frame  kernel {d['roofline']['kernel_ms_per_frame']:8.3f} ms  parity {d.get('parity', {}).get('ok')}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(v, "FAILED", e, open(f"gpurun_out/ab_{v}.err").read()[-400:])
PY
done
