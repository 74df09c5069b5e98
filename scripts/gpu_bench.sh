#!/bin/bash
# smoke + default bench + ncu launch list on the GPU box; logs into gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
fi
