#!/bin/bash
# decoder iteration: speller parity tests, then the c2 bench (phase timers of the folded decoder in bench.err)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_speller.py -m gpu -q -x --tb=short -p no:cacheprovider ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/test_gpu_speller.log 2>&1
echo "speller rc=$?" | tee -a gpurun_out/summary.txt
tail -n 25 gpurun_out/test_gpu_speller.log
PLAS_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err; echo "bench(dbg) rc=$?" | tee -a gpurun_out/summary.txt
grep -E "decoder fold|decoder tc" gpurun_out/bench_dbg.err | tail -n 6
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]))
    for k,v in d["stages"].items(): print(f"  {k:12s} {v['ms_per_step']:8.3f} ms  frac={v.get('frac')}")
except Exception as e: print("no bench json", e)
PY
