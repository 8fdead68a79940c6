#!/bin/bash
# A/B bench of library variants built with extra -D flags (see raymarchcl_b200/csrc/build.py -o):
#   tools/gpu_ab.sh name1 name2 ...   runs bench.py (resident arm only) with RAYMARCH_B200_LIB=build_ab/<name>.so
set -u
mkdir -p gpurun_out
for v in "$@"; do
  RAYMARCH_B200_LIB=$PWD/build_ab/$v.so timeout 300 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} \
    > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{v}.json"))
    print(f"{v:24s} {d['ms_per_step']:8.3f} ms/frame  kernel {d['roofline']['kernel_ms_per_frame']:8.3f} ms  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(v, "FAILED", e, open(f"gpurun_out/ab_{v}.err").read()[-400:])
PY
done
