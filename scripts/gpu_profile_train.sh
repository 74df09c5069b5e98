#!/bin/bash
# GPU evidence run for the TRAINING step (c3): smoke, bench line, ncu launch list, ncu --set full of the top kernels.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --workload c3 --impl reference --steps 2 --warmup 0 > gpurun_out/bench_c3_ref.json 2> gpurun_out/bench_c3_ref.err; echo "bench c3 ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
    python scripts/train_probe.py --one > gpurun_out/ncu_train.log 2>&1; echo "ncu launches rc=$?"
# kernel, launch-skip (a representative launch: layer-1 projection GEMM, layer-0 recurrences, a mid-sequence decoder step)
for spec in gemm_tf32x3_tcgen05_kernel:2 gemm_f32_ex_kernel:2 rec_train_fwd_kernel:0 rec_train_bwd_kernel:2 dec_cell_fwd_kernel:20 dec_gemv_t_kernel:20 dec_att_fwd_kernel:20 dec_att_bwd_kernel:20; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/prof_train_$k \
      python scripts/train_probe.py --one > gpurun_out/ncu_train_$k.log 2>&1; echo "ncu $k rc=$?"
done
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3.json'))
print("c3 value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d["stages"].items(): print(f"  {k:14s} {v['ms_per_step']:8.3f} ms")
PY
