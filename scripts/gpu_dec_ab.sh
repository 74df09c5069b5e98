#!/bin/bash
# decoder A/B: speller parity tests, then the c2 decoder stage time with and without multicast activation loads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_speller.py tests/test_gpu_baseline_shapes.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/test_gpu_speller.log 2>&1
echo "speller rc=$?"; tail -n 4 gpurun_out/test_gpu_speller.log
for mc in 1 1; do
  PLAS_DEC_MCAST=$mc PLAS_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --streams 1 > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err
  echo "mcast=$mc"; grep -E "decoder fold phase|gemm phase 0 detail" gpurun_out/bench_dbg.err | tail -n 2
  PLAS_DEC_MCAST=$mc python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  value', round(d['value']), 'e2e', round(d['e2e']['value']), 'decoder', round(d['stages']['decoder']['ms_per_step'],3))"
done
