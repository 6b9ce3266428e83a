#!/bin/bash
# Round 2, third pass: full ncu captures of the kernels that changed (run under gpurun, ONE GPU):
#   gpurun --timeout 900 -- 'bash tools/gpu_profile_pass3.sh'
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
S="python tools/bench_samplers.py --iters 1 --warmup 1 --what"
timeout 120 $S importance packed > gpurun_out/prof3_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof3_plain.log; exit 1; }
timeout 300 $N -k regex:"coarse_fwd_packed|stream_packed|bwd_span_packed" -c 6 -o gpurun_out/r02ah_packed_composite $S packed > gpurun_out/ncu_ah1.log 2>&1; echo "packed composite rc=$?"
timeout 300 $N -k regex:importance_grp_kernel -s 4 -c 4 -o gpurun_out/r02ah_imp_packed $S packed > gpurun_out/ncu_ah2.log 2>&1; echo "packed importance rc=$?"
timeout 300 $N -k regex:importance_grp_kernel -c 8 -o gpurun_out/r02ah_imp_dense $S importance > gpurun_out/ncu_ah3.log 2>&1; echo "dense importance rc=$?"
timeout 200 $N -k regex:composite_fwd_span_kernel -s 2 -c 2 -o gpurun_out/r02ah_dense_pipeline python tools/bench_dense_pipeline.py --iters 1 > gpurun_out/ncu_ah4.log 2>&1; echo "dense pipeline rc=$?"
ls -la gpurun_out/r02ah_*.ncu-rep
