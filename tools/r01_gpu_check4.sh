#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_field.py -q --tb=short -p no:cacheprovider > gpurun_out/r01e_field_tests.log 2>&1; echo "field pytest rc=$?"; tail -2 gpurun_out/r01e_field_tests.log
AVR_FIELD_SHARE_POINT=1 timeout 50 python tools/bench_field.py --iters 5 --raw-only > gpurun_out/r01e_field_share.jsonl 2>&1; echo "share rc=$?"
AVR_FIELD_SHARE_POINT=0 timeout 50 python tools/bench_field.py --iters 5 --raw-only > gpurun_out/r01e_field_noshare.jsonl 2>&1; echo "noshare rc=$?"
grep -h "^{" gpurun_out/r01e_field_share.jsonl gpurun_out/r01e_field_noshare.jsonl | cut -c1-200
timeout 150 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_field.py > gpurun_out/r01e_all_tests.log 2>&1; echo "full pytest rc=$?"; tail -4 gpurun_out/r01e_all_tests.log
