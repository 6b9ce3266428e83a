#!/bin/bash
mkdir -p gpurun_out
timeout 80 python tools/bench_field.py --iters 5 > gpurun_out/r01c_field.jsonl 2>&1; echo "field rc=$?"
AVR_FIELD_NOCACHE=1 timeout 60 python tools/bench_field.py --iters 5 --raw-only > gpurun_out/r01c_field_nocache.jsonl 2>&1; echo "nocache rc=$?"
AVR_FIELD_BWD_SPLIT=1 timeout 60 python tools/bench_field.py --iters 5 --raw-only > gpurun_out/r01c_field_split.jsonl 2>&1; echo "split rc=$?"
timeout 60 python tools/bench_field.py --iters 3 --raw-only --rays 8192 --batches 3 > gpurun_out/r01c_field_big.jsonl 2>&1; echo "big rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:field_inputs -c 4 -f -o gpurun_out/r01c_field python tools/bench_field.py --iters 1 --batches 1 --raw-only > gpurun_out/r01c_ncu.log 2>&1; echo "ncu rc=$?"
grep -h "^{" gpurun_out/r01c_field.jsonl gpurun_out/r01c_field_nocache.jsonl gpurun_out/r01c_field_split.jsonl gpurun_out/r01c_field_big.jsonl | cut -c1-260
