#!/bin/bash
# phase timers of the decoder on the c2 bench shape (PLAS_DEBUG synchronises: not a bench number)
mkdir -p gpurun_out
PLAS_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err; echo "bench(dbg) rc=$?"
grep -E "decoder fold|decoder tc|gemm phase|\[plas\]" gpurun_out/bench_dbg.err | tail -n ${TAIL:-4}
