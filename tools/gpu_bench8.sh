#!/bin/bash
# 8-GPU bench lines (one torchrun launch per workload)
set -u
mkdir -p gpurun_out
N=${N:-8}
for wl in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --workload $wl > gpurun_out/bench_${wl}_${N}gpu.json 2> gpurun_out/bench_${wl}_${N}gpu.err
  echo "$wl rc=$?"; cut -c1-260 gpurun_out/bench_${wl}_${N}gpu.json; tail -c 300 gpurun_out/bench_${wl}_${N}gpu.err
done
