#!/bin/bash
# compute-sanitizer over the bf16 tensor-core kernels (rec_tc, decoder_fold, gemm_bf16, gemm_tf32x3, front-end): memcheck and
# synccheck (barrier / mbarrier misuse), on the small parity cases (the tools slow kernels down 10-50x).
mkdir -p gpurun_out
K='(greedy_parity and bf16) or (test_listener_parity and bf16) or test_teacher_forced_parity or tf32x3 or (test_frontend_parity and w25)'
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_gpu_speller.py tests/test_gpu_listener.py tests/test_gpu_gemm_tf32.py tests/test_gpu_frontend.py -m gpu -q -p no:cacheprovider -k "$K" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -E "passed|failed|ERROR SUMMARY|SUMMARY" gpurun_out/sanitizer_$tool.log | tail -4
done
# the BASELINE-width decoder (128 CTAs in clusters of 4, resident keys, bulk-copied activation tiles) and the shared-memory
# hazards of the folded decoder's phase-shared regions (racecheck, small cases)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_speller.py -m gpu -q -p no:cacheprovider -k "full_width" > gpurun_out/sanitizer_memcheck_fullwidth.log 2>&1
echo "memcheck(full width) rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_memcheck_fullwidth.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_speller.py -m gpu -q -p no:cacheprovider -k "greedy_parity and bf16 and (B5-Tm19 or B33 or B9-Tm37 or B40)" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_racecheck.log | tail -4
