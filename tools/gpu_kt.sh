#!/bin/bash
# developed-state kernel times only (fast iteration during kernel optimisation); optional parity tests with $1=test
mkdir -p gpurun_out
if [ "$1" = "test" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
timeout 600 python tools/prof_run.py --ncrms 47360 --nstop 10 --calls 2 --spinup ${SPIN:-60} --ktimes 2>&1 | tee gpurun_out/ktimes_dev.log | head -16
