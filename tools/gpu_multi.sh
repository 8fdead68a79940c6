#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): group-context tests, torchrun bench lines, single-process group bench.
set -u
mkdir -p gpurun_out
N=${N:-2}
TAG=${1:-r02_multi}
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_configs.py::test_default_kernel_equals_round1_kernel_bit_for_bit" -x -q > gpurun_out/${TAG}_${N}gpu_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_${N}gpu_pytest.log
for wl in ${WORKLOADS:-c2}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --workload $wl > gpurun_out/${TAG}_bench_${wl}_${N}gpu.json 2> gpurun_out/${TAG}_bench_${wl}_${N}gpu.err
  echo "$wl torchrun rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_${wl}_${N}gpu.json"))
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["per_rank"], d["e2e"]["ms_per_step"] if d.get("e2e") else None, d["parity"])
except Exception as e:
    print("no line:", e)
PY
  tail -c 400 gpurun_out/${TAG}_bench_${wl}_${N}gpu.err
done
GL=1; for k in 2 4 8; do [ $k -le $N ] && GL=$GL,$k; done
timeout 900 python tools/bench_multi.py --gpus $GL --steps ${STEPS:-10} > gpurun_out/${TAG}_group_${N}gpu.jsonl 2> gpurun_out/${TAG}_group_${N}gpu.err
echo "group bench rc=$?"; cat gpurun_out/${TAG}_group_${N}gpu.jsonl; tail -c 400 gpurun_out/${TAG}_group_${N}gpu.err
