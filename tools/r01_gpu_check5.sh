#!/bin/bash
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_field.py -q --tb=short -p no:cacheprovider > gpurun_out/r01f_field_tests.log 2>&1; echo "field pytest rc=$?"; tail -2 gpurun_out/r01f_field_tests.log
AVR_FIELD_STAGE=1 timeout 50 python tools/bench_field.py --iters 5 --raw-only > gpurun_out/r01f_field_stage.jsonl 2>&1; echo "stage rc=$?"
AVR_FIELD_STAGE=0 timeout 50 python tools/bench_field.py --iters 5 --raw-only > gpurun_out/r01f_field_nostage.jsonl 2>&1; echo "nostage rc=$?"
AVR_FIELD_STAGE=1 timeout 50 python tools/bench_field.py --iters 3 --raw-only --rays 8192 --batches 3 > gpurun_out/r01f_field_stage_big.jsonl 2>&1; echo "stage big rc=$?"
grep -h "^{" gpurun_out/r01f_field_stage.jsonl gpurun_out/r01f_field_nostage.jsonl gpurun_out/r01f_field_stage_big.jsonl | cut -c1-200
