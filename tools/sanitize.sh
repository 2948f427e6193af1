#!/bin/bash
# compute-sanitizer passes over the small GPU parity tests (run on the GPU box; results under gpurun_out/).
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards; synccheck: barrier misuse;
# initcheck: reads of uninitialised device memory.
set -u
mkdir -p gpurun_out
SEL='not full_size and not 4096 and not cfg2 and not cfg1 and not device_srs'
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_lasso.py tests/test_gpu_kzg.py tests/test_gpu_generic.py \
    tests/test_gpu_hyperplonk.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_summary.txt
  grep -c "=========" gpurun_out/sanitize_$tool.log | tee -a gpurun_out/sanitize_summary.txt
  tail -3 gpurun_out/sanitize_$tool.log
done
