#!/bin/bash
# One gpurun call: GPU parity tests, a short bench, kernel launch list and one full ncu capture of a steady sub-step.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_run.py --ncrms 47360 --nstop 6 --calls 3 --ktimes > gpurun_out/ktimes.log 2>&1
cat gpurun_out/ktimes.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python tools/prof_run.py --ncrms 4736 --nstop 3 --calls 2 > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(p[0-9]|mpdata|precip_[ab]|stats|kurant)' -s 40 -c 16 \
   -o gpurun_out/prof_r1 -f python tools/prof_run.py --ncrms 4736 --nstop 6 --calls 1 > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log
timeout 1500 python bench.py --steps 1 --warmup 1 --nstop 60 > gpurun_out/bench_short.log 2>&1
tail -2 gpurun_out/bench_short.log
