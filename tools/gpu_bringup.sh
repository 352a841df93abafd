#!/bin/bash
# Run every GPU test in its own process with a timeout, so a trapped kernel cannot take the
# other tests (or the box) with it.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
python -m feature_vs_text_compound_emotion_b200.build || exit 1
FILES="${@:-tests/test_gpu_kernels.py tests/test_gpu_lfan.py}"
: > gpurun_out/bringup.log
pass=0; fail=0
for id in $(python -m pytest $FILES --collect-only -q -m gpu 2>/dev/null | grep "::"); do
  if timeout 300 python -m pytest "$id" -x -q -m gpu > gpurun_out/one.log 2>&1; then
    pass=$((pass+1)); echo "PASS $id" | tee -a gpurun_out/bringup.log
  else
    fail=$((fail+1)); echo "FAIL $id" | tee -a gpurun_out/bringup.log
    grep -E "Error|error|assert|watchdog|relative|cos" gpurun_out/one.log | head -12 | tee -a gpurun_out/bringup.log
  fi
done
echo "passed=$pass failed=$fail" | tee -a gpurun_out/bringup.log
