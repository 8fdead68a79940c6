#!/bin/bash
# Final single-GPU bench lines of the round: headline (both arms) + the other BASELINE configs that fit one GPU.
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/final_c2_1gpu.json 2> gpurun_out/final_c2_1gpu.err; echo "c2 rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/final_c2_reference_arm.json 2> gpurun_out/final_c2_reference_arm.err; echo "ref rc=$?"
for wl in c1 c3 c5; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_${wl}_1gpu.json 2> gpurun_out/final_${wl}_1gpu.err; echo "$wl rc=$?"
done
for f in gpurun_out/final_*.json; do echo "$f: $(cut -c1-220 $f)"; done
