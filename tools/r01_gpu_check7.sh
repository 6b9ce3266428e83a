#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_field.py -q --tb=short -p no:cacheprovider > gpurun_out/r01h_field_tests.log 2>&1; echo "field pytest rc=$?"; tail -2 gpurun_out/r01h_field_tests.log
AVR_FIELD_SHARE_POINT=1 AVR_FIELD_BWD_PREFETCH=1 timeout 40 python tools/bench_field.py --iters 5 --raw-only > gpurun_out/r01h_field_prefetch.jsonl 2>&1; echo "prefetch rc=$?"
timeout 40 python tools/bench_field.py --iters 5 --raw-only > gpurun_out/r01h_field_default.jsonl 2>&1; echo "default rc=$?"
grep -h "^{" gpurun_out/r01h_field_prefetch.jsonl gpurun_out/r01h_field_default.jsonl | cut -c1-200
