#!/bin/bash
# Full GPU evidence run: parity suite, bench, ncu launch list, ncu --set full captures of the top kernels.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -n 6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
for k in rec_tc_kernel decoder_tc_kernel gemm_bf16_tcgen05_kernel fe_spectral_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?" | tee -a gpurun_out/summary.txt
done
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d["stages"].items(): print(f"  {k:12s} {v['ms_per_step']:8.3f} ms  frac={v.get('frac')}")
PY
