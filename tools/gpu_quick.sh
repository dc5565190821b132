#!/bin/bash
# quick GPU check: parity tests + per-kernel times at a near-full batch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_run.py --ncrms 47360 --nstop 6 --calls 3 --ktimes > gpurun_out/ktimes.log 2>&1
head -20 gpurun_out/ktimes.log
timeout 600 python tools/prof_run.py --ncrms 9472 --spinup 180 --nstop 10 --calls 2 --ktimes > gpurun_out/ktimes_dev.log 2>&1
head -20 gpurun_out/ktimes_dev.log
