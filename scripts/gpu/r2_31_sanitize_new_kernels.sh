#!/bin/bash
# round 2, call 31: compute-sanitizer on the kernels added in the last session: GRU cluster kernels (H = 256, 40 rows = 2 full 16-row clusters + a
# ragged one) and the CTA-pair GEMM inside an H = 512 step (dU / dW: M 512, N 2048, split-K accumulate -> pair form)
set -x
mkdir -p gpurun_out
python scripts/sanitize_case.py 256 8 40 bf16 GRU
python scripts/sanitize_case.py 512 8 72 bf16 LSTM
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 6 python scripts/sanitize_case.py 256 8 40 bf16 GRU > gpurun_out/r2_31_san_${tool}_gru256.log 2>&1
  echo "$tool gru rc=$?"; grep -E "SUMMARY|sanitize_case" gpurun_out/r2_31_san_${tool}_gru256.log | tail -3
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 6 python scripts/sanitize_case.py 512 8 72 bf16 LSTM > gpurun_out/r2_31_san_memcheck_512_pairgemm.log 2>&1
echo "memcheck 512 rc=$?"; grep -E "SUMMARY|sanitize_case" gpurun_out/r2_31_san_memcheck_512_pairgemm.log | tail -3
