#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r01g_all_tests.log 2>&1; echo "full pytest rc=$?"; tail -3 gpurun_out/r01g_all_tests.log
timeout 60 python tools/bench_field.py --iters 5 > gpurun_out/r01g_field.jsonl 2>&1; echo "field rc=$?"
grep -h "^{" gpurun_out/r01g_field.jsonl | cut -c1-330
timeout 100 ncu --set full --clock-control none --import-source on -k regex:field_inputs -c 3 -f -o gpurun_out/r01g_field python tools/bench_field.py --iters 1 --batches 1 --raw-only > gpurun_out/r01g_ncu.log 2>&1; echo "ncu rc=$?"
