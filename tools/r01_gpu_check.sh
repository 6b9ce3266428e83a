#!/bin/bash
# One bounded GPU call: parity of the new kernels (packed coarse sampler, radiance-field front end),
# then their measurements.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_field.py tests/test_gpu_packed.py -q --tb=short -p no:cacheprovider > gpurun_out/r01b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r01b_tests.log
tail -15 gpurun_out/r01b_tests.log
timeout 70 python tools/bench_field.py --iters 5 > gpurun_out/r01b_field.jsonl 2>&1; echo "field rc=$?"
AVR_COARSE_PACKED=ray timeout 60 python tools/bench_packed_pipeline.py --iters 3 > gpurun_out/r01b_c4_ray.jsonl 2>&1; echo "c4 ray rc=$?"
timeout 60 python tools/bench_packed_pipeline.py --iters 3 > gpurun_out/r01b_c4_flat.jsonl 2>&1; echo "c4 flat rc=$?"
tail -4 gpurun_out/r01b_field.jsonl
head -1 gpurun_out/r01b_c4_ray.jsonl; head -1 gpurun_out/r01b_c4_flat.jsonl; tail -1 gpurun_out/r01b_c4_flat.jsonl
