#!/bin/bash
# Full GPU evidence run: parity suite, bench, ncu launch list, ncu --set full captures of the top kernels.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -n 6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
if [ "${REF:-1}" = "1" ]; then
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_ref.json
fi
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    $B > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
# one full capture per kernel family.  -s skips the warm-up launches of that kernel: rec / fe launch 4 resp. 1 times per pass
# (capture the layer-0 launch of pass 2 / the 2nd pass); the GEMM capture takes the FOUR in-projection launches of one pass
# (8 GEMM launches per pass: 4 projections, keys, VW, PV ...: -s 7 starts at the second pass' layer 0)
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rec_tc_kernel -s 4 -c 1 -f -o gpurun_out/prof_rec_tc_kernel $B1 > gpurun_out/ncu_rec_tc_kernel.log 2>&1; echo "ncu rec rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_fold_kernel -s 1 -c 1 -f -o gpurun_out/prof_decoder_fold_kernel $B1 > gpurun_out/ncu_decoder_fold_kernel.log 2>&1; echo "ncu decoder rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_kernel -s 7 -c 4 -f -o gpurun_out/prof_gemm_bf16_tcgen05_kernel $B1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fe_spectral_kernel -s 1 -c 1 -f -o gpurun_out/prof_fe_spectral_kernel $B1 > gpurun_out/ncu_fe.log 2>&1; echo "ncu fe rc=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d["stages"].items(): print(f"  {k:12s} {v['ms_per_step']:8.3f} ms  frac={v.get('frac')}")
PY
