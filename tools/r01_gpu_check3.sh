#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_field.py -q --tb=short -p no:cacheprovider > gpurun_out/r01d_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r01d_tests.log
timeout 80 python tools/bench_field.py --iters 5 > gpurun_out/r01d_field.jsonl 2>&1; echo "field rc=$?"
timeout 60 python tools/bench_field.py --iters 3 --raw-only --rays 8192 --batches 3 > gpurun_out/r01d_field_big.jsonl 2>&1; echo "big rc=$?"
grep -h "^{" gpurun_out/r01d_field.jsonl gpurun_out/r01d_field_big.jsonl | cut -c1-300
