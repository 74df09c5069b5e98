#!/bin/bash
# A/B of the device-resident arm's pacing: batches the host keeps in flight in the serving loop (PLAS_BENCH_AHEAD)
mkdir -p gpurun_out; : > gpurun_out/ahead.txt
for rep in 1 2 3; do for a in 2 3; do
  PLAS_BENCH_AHEAD=$a timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/ahead_$a.json 2> gpurun_out/ahead_$a.err
  python - $a <<'PY' | tee -a gpurun_out/ahead.txt
import json,sys
d=json.loads(open(f'gpurun_out/ahead_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print("ahead", sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
PY
done; done
