#!/bin/bash
# quick iteration: selected GPU tests + bench (no ncu).  usage: gpu_iter.sh <pytest -k expr or ''> [test files...]
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q -x --tb=short -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "$name rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 15 "gpurun_out/$name.log"
done
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]))
    for k,v in d["stages"].items(): print(f"  {k:12s} {v['ms_per_step']:8.3f} ms  frac={v.get('frac')}")
except Exception as e: print("no bench json", e)
PY
