#!/bin/bash
# Full ncu captures of the sampling kernels (run under gpurun, ONE GPU):
#   gpurun --timeout 900 -- 'bash tools/gpu_profile_samplers.sh'
# importance 64->128 with merge, 64->16+16 with merge, the four packed classes, the packed coarse sampler.
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
S="python tools/bench_samplers.py --iters 1 --warmup 1 --what"
timeout 120 $S importance packed > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
timeout 200 $N -k regex:importance_grp_kernel -s 1 -c 1 -o gpurun_out/r02q_imp_c3 $S importance > gpurun_out/ncu_q1.log 2>&1; echo "c3 rc=$?"
timeout 200 $N -k regex:importance_grp_kernel -s 5 -c 1 -o gpurun_out/r02q_imp_default $S importance > gpurun_out/ncu_q2.log 2>&1; echo "default rc=$?"
timeout 300 $N -k regex:importance_grp_kernel -s 4 -c 4 -o gpurun_out/r02q_imp_packed $S packed > gpurun_out/ncu_q3.log 2>&1; echo "packed rc=$?"
timeout 200 $N -k regex:coarse_fwd_packed -s 1 -c 1 -o gpurun_out/r02q_coarse_packed $S packed > gpurun_out/ncu_q4.log 2>&1; echo "coarse packed rc=$?"
ls -la gpurun_out/r02q_*.ncu-rep
