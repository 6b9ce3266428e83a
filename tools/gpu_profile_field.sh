#!/bin/bash
# Full ncu captures of the radiance-field front end (forward, both backward launches) and of the LSTM march
# (run under gpurun, ONE GPU):   gpurun --timeout 900 -- 'bash tools/gpu_profile_field.sh'
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 120 python tools/bench_field.py --iters 2 --raw-only > gpurun_out/prof_field_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_field_plain.log; exit 1; }
timeout 300 $N -k regex:field_inputs -s 6 -c 6 -o gpurun_out/r02r_field python tools/bench_field.py --iters 2 --raw-only > gpurun_out/ncu_r1.log 2>&1; echo "field rc=$?"
timeout 300 $N -k regex:lstm_march -s 2 -c 4 -o gpurun_out/r02r_march python tools/bench_march.py --iters 1 > gpurun_out/ncu_r2.log 2>&1; echo "march rc=$?"
ls -la gpurun_out/r02r_*.ncu-rep
