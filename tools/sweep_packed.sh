#!/bin/bash
# Tuning sweep of the packed span kernels' compile-time knobs (samples per lane, warps per CTA, ring slots):
# rebuilds the library on the GPU box per variant and times tools/bench_samplers.py --what packed.
# Usage (under gpurun): [SWEEP="13 2 2,13 2 3"] bash tools/sweep_packed.sh > gpurun_out/sweep_packed.log
IFS=","
for v in ${SWEEP:-13 2 2,13 4 2,11 2 2,13 2 3}; do
  IFS=" "
  set -- $v
  AVR_NVCC_EXTRA="-DAVR_PK_L=$1 -DAVR_PK_WARPS=$2 -DAVR_PK_STAGES=$3" python adaptive-volume-rendering_b200/build.py --force > /dev/null 2>&1 || { echo "build failed L=$1 W=$2 NS=$3"; continue; }
  echo "== L=$1 warps=$2 slots=$3"
  python tools/bench_samplers.py --what packed 2>&1 | grep -E "composite_(fwd|bwd)_packed" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', d['kernel'], d['ms'], d['hbm_frac'])"
done
python adaptive-volume-rendering_b200/build.py --force > /dev/null 2>&1
