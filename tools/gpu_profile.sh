#!/bin/bash
# One GPU-box visit: parity tests, the headline bench line, the ncu launch list of the same command
# and one `ncu --set full` capture of the render kernel (reduced frame), all into gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-run}
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_c2.json
[ "${SKIP_NCU:-0}" = "1" ] && exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e \
  > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_bricks -s 1 -c 1 -f \
  -o gpurun_out/${TAG}_bricks_c2s python bench.py --workload c2s --steps 1 --warmup 1 --no-cpu-baseline --no-e2e \
  > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out
if [ "${FULL_C2:-0}" = "1" ]; then
  # full-size capture of one render launch (all 16 passes of the 1080p frame): dram__bytes for roofline.traffic
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_bricks -s 1 -c 1 -f \
    -o gpurun_out/${TAG}_bricks_c2 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e \
    > gpurun_out/${TAG}_ncu_full_c2.log 2>&1
  echo "ncu full c2 rc=$?"
fi
