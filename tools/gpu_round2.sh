#!/bin/bash
# One GPU-box visit of round 2: parity tests, headline bench (default kernel), A/B of the block sizes and of
# round 1's kernel, ncu launch list + one full capture of the default kernel. Everything into gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
nvidia-smi -L; nproc
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench_c2.json; tail -3 gpurun_out/${TAG}_bench_c2.err
if [ "${SKIP_AB:-0}" != "1" ]; then
  for B in 512 768 1024; do
    timeout 300 python bench.py --steps 10 --warmup 3 --opt 10=$B --no-e2e --no-cpu-baseline --no-parity \
      > gpurun_out/${TAG}_ab_block$B.json 2> gpurun_out/${TAG}_ab_block$B.err
    echo "block $B: $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_ab_block$B.json'));print(d['ms_per_step'], d['roofline']['kernel_ms_per_frame'])")"
  done
  timeout 300 python bench.py --steps 10 --warmup 3 --kernel bricks --no-e2e --no-cpu-baseline --no-parity \
    > gpurun_out/${TAG}_ab_bricks.json 2> gpurun_out/${TAG}_ab_bricks.err
  echo "bricks: $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_ab_bricks.json'));print(d['ms_per_step'], d['roofline']['kernel_ms_per_frame'])")"
fi
[ "${SKIP_NCU:-0}" = "1" ] && exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity \
  > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"
WL=${NCU_WORKLOAD:-c2}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_render_persist -s 1 -c 1 -f \
  -o gpurun_out/${TAG}_persist_${WL} python bench.py --workload ${WL} --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity ${NCU_OPTS:-} \
  > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out
