#!/bin/bash
# Round-end captures (run under gpurun, ONE GPU): the launch list of the headline bench and full ncu captures of
# the span kernels and of the importance sampler as shipped.   gpurun --timeout 900 -- 'bash tools/gpu_profile_final.sh'
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu-baseline"
timeout 120 $B > gpurun_out/final_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/final_plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_final.csv $B > gpurun_out/ncu_launches_final.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:composite_.*_span_kernel -s 6 -c 2 -o gpurun_out/r02_span_final $B > gpurun_out/ncu_span_final.log 2>&1; echo "span capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:importance_grp_kernel -s 1 -c 1 -o gpurun_out/r02_imp_c3_final python tools/bench_samplers.py --what importance --iters 1 --warmup 1 > gpurun_out/ncu_imp_final.log 2>&1; echo "importance capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:field_inputs_bwd_latent_async -s 2 -c 1 -o gpurun_out/r02_field_latent_final python tools/bench_field.py --iters 1 --raw-only > gpurun_out/ncu_field_final.log 2>&1; echo "field capture rc=$?"
ls -la gpurun_out/*final*.ncu-rep
