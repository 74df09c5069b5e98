#!/bin/bash
# Run the GPU parity suites one file at a time (each under its own timeout) and keep the logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu.txt
status=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q -x --tb=short -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 25 "gpurun_out/$name.log"
  [ $rc -ne 0 ] && status=1
done
exit $status
