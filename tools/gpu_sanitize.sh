#!/bin/bash
# compute-sanitizer over the small configuration for every render kernel. Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_c1.py ${KERNELS:-0 1 3 4} \
    > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_${tool}.log
done
