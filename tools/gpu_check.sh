#!/bin/bash
# One bounded GPU call (run under gpurun): parity first, then the measurement tools.
#   gpurun --timeout 400 -- 'bash tools/gpu_check.sh'
# Everything lands in gpurun_out/ (scratch); copy what should be judged into profiles/.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 80 python tools/bench_field.py --iters 5 > gpurun_out/field.jsonl 2>&1; echo "bench_field rc=$?"
timeout 80 python tools/bench_packed_pipeline.py --iters 3 > gpurun_out/c4_pipeline.jsonl 2>&1; echo "bench_packed_pipeline rc=$?"
grep -h "^{" gpurun_out/field.jsonl gpurun_out/c4_pipeline.jsonl | cut -c1-300
