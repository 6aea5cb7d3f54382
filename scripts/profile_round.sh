#!/bin/bash
# End-of-round evidence on one B200 (run through gpurun): ncu metrics of the hot kernels at config-2 shapes, an ncu launch
# list of bench.py without the CUDA graph (cut ONE training step out of it: profiles/README.md), then the default bench line.
# Each command runs plainly first; ncu only after it exited 0.  The launch list takes ~10 minutes under ncu.
mkdir -p gpurun_out/logs gpurun_out/ncu
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__inst_executed_pipe_tensor.sum
timeout 120 python scripts/prof_small.py > gpurun_out/logs/prof_small_plain.log 2>&1 && \
timeout 400 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu/r2b_hot_kernels_raw.csv python scripts/prof_small.py > gpurun_out/logs/prof_small_ncu.log 2>&1
echo "hot kernels rc=$?"
timeout 200 python bench.py --steps 1 --warmup 3 --no-graph --no-gen --no-cpu --no-ref-gpu > gpurun_out/logs/bench_nograph_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu/r2b_train_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-gen --no-cpu --no-ref-gpu > gpurun_out/logs/bench_nograph_ncu.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/ncu/
timeout 500 python bench.py > gpurun_out/logs/bench_final_n1.json 2> gpurun_out/logs/bench_final_n1.err
echo "bench rc=$?"; head -c 300 gpurun_out/logs/bench_final_n1.json
