#!/bin/bash
# One bounded profiling call (run under gpurun, ONE GPU): launch list of the headline bench, full ncu
# captures of the span kernels, the importance sampler and the LSTM march, and the host-pipeline sweep.
#   gpurun --timeout 900 -- 'bash tools/gpu_profile.sh'
# Everything lands in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu-baseline"
timeout 120 $B > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:composite_.*_span_kernel -s 6 -c 2 -o gpurun_out/r02_span $B > gpurun_out/ncu_span.log 2>&1; echo "span capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:importance_grp_kernel -s 3 -c 1 -o gpurun_out/r02_importance python tools/bench_samplers.py --what importance --iters 1 --warmup 1 > gpurun_out/ncu_imp.log 2>&1; echo "importance capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_march -s 2 -c 2 -o gpurun_out/r02_march python tools/bench_march.py --iters 1 > gpurun_out/ncu_march.log 2>&1; echo "march capture rc=$?"
timeout 200 python tools/sweep_host_chunks.py > gpurun_out/r02_host_sweep.jsonl 2>&1; echo "host sweep rc=$?"; cat gpurun_out/r02_host_sweep.jsonl | cut -c1-120
timeout 120 python tools/bench_dropin.py --iters 10 > gpurun_out/r02_dropin.json 2>gpurun_out/r02_dropin.err; echo "dropin rc=$?"
ls -la gpurun_out/*.ncu-rep
