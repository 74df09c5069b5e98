#!/bin/bash
# front-end iteration: its parity tests, the c5 sweep line and the executed-instruction count of fe_spectral_kernel
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_frontend.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/fe_tests.log 2>&1; echo "fe tests rc=$?" | tee -a gpurun_out/summary.txt
tail -n 8 gpurun_out/fe_tests.log
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1])
print("c5 value", d.get("value"), "ms", d.get("ms_per_step"))
print(json.dumps(d.get("variants", d.get("config")), indent=0)[:1500])
PY
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:fe_spectral_kernel -s 1 -c 1 --csv --log-file gpurun_out/fe_inst.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/fe_inst.log 2>&1; echo "ncu fe rc=$?" | tee -a gpurun_out/summary.txt
grep -E "inst_executed|time_duration" gpurun_out/fe_inst.csv | tail -n 4
