#!/bin/bash
# ncu launch list of the DEFAULT bench command (python bench.py), taken inside its timed region: the first 9000 kernel
# launches (three warm-up calls of ~2800 launches) are skipped, the next 1400 (half a timed call, ~90 CRM steps) listed.
# Per-launch times under ncu are serialised/cold-cache: only the kernels' SHARES of the step are compared with bench.py.
mkdir -p gpurun_out
timeout 1800 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 1400 --csv --log-file gpurun_out/launches_default.csv \
   python bench.py --no-cpu --e2e-steps 0 > gpurun_out/launches_default.log 2>&1
tail -1 gpurun_out/launches_default.log | cut -c1-300
