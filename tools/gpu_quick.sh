#!/bin/bash
# Quick gpurun call: GPU parity tests + developed-state kernel times (+ optional extra command in $1)
set -x
mkdir -p gpurun_out
(which gfortran flang nvfortran ifort; nproc; lscpu | grep "Model name") > gpurun_out/host.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/prof_run.py --ncrms 47360 --nstop 10 --calls 2 --spinup 180 --ktimes > gpurun_out/ktimes_dev.log 2>&1
cat gpurun_out/ktimes_dev.log
if [ -n "$1" ]; then eval "$1"; fi
