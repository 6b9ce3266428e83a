#!/bin/bash
# Limiter survey (run under gpurun, ONE GPU): DRAM / LSU-pipe / issue utilisation of every kernel of this library launched
# by the dense pipeline, the sampler microbenchmarks and the packed pipeline.   gpurun --timeout 900 -- 'bash tools/gpu_limiter_survey.sh'
mkdir -p gpurun_out
M="gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,launch__registers_per_thread"
K='regex:composite|importance|ray_points|coarse|sort_rays|rays_|depth_from|world_rays|field_inputs|lstm'
timeout 250 ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/survey_dense.csv python tools/bench_dense_pipeline.py --iters 1 > gpurun_out/survey_dense.log 2>&1; echo "dense rc=$?"
timeout 250 ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/survey_samplers.csv python tools/bench_samplers.py --iters 1 --warmup 1 --what coarse geometry adaptive > gpurun_out/survey_samplers.log 2>&1; echo "samplers rc=$?"
timeout 300 ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/survey_packed.csv python tools/bench_packed_pipeline.py --rays 1048576 --iters 1 > gpurun_out/survey_packed.log 2>&1; echo "packed rc=$?"
wc -l gpurun_out/survey_*.csv
