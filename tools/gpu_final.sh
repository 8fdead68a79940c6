#!/bin/bash
# Round-end visit: all GPU tests, the headline bench line + reference arm, launch list, full ncu captures.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
cp raymarchcl_b200/libraymarch_b200.so gpurun_out/${TAG}_lib.so
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_c2_1gpu.json 2> gpurun_out/${TAG}_bench_c2_1gpu.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_c2_1gpu.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_c2_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err
echo "reference arm rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_c2_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
for wl in c1 c3 c5; do
  timeout 900 python bench.py --workload $wl --steps 10 > gpurun_out/${TAG}_bench_${wl}_1gpu.json 2> gpurun_out/${TAG}_bench_${wl}_1gpu.err
  echo "$wl rc=$? $(cut -c1-160 gpurun_out/${TAG}_bench_${wl}_1gpu.json)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity \
  > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"
ncu_cap() {  # name, extra bench args
  local name=$1; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_persist -s 1 -c 1 -f \
    -o gpurun_out/${TAG}_ncu_${name} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity "$@" \
    > gpurun_out/${TAG}_ncu_${name}.log 2>&1
  echo "ncu $name rc=$?"
}
ncu_cap final_c2                                              # the default: 1024 x 1, free-running warps, distance map staged into shared memory by TMA
ncu_cap final_c2_256x5 --opt 10=256 --opt 12=0                # 256 x 5, free-running, byte map in global memory (the default of short launches)
ncu_cap final_c2_rounds --opt 10=256 --opt 12=0 --opt 11=1    # ... with block-synchronous rounds
ls -la gpurun_out | tail -30
