#!/bin/bash
# full bench (both arms) + ncu launch list of the bench command
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
timeout 1500 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --nstop 8 --no-cpu --e2e-steps 0 > gpurun_out/launches.log 2>&1
tail -2 gpurun_out/launches.log
