#!/bin/bash
# DRAM bytes per launch of every kernel of one developed-state sub-step at the FULL cfg2 batch (48,602 CRMs): one
# single-pass ncu capture (no kernel replay, so nothing has to be saved/restored around the 8 ms launches).
# The profiler range starts after the untimed spin-up call (tools/prof_run.py --spinup -> cudaProfilerStart).
mkdir -p gpurun_out
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
   -k regex:'k_(p[0-9]|mpdata|precip_[ab]|stats|kurant)' -c ${1:-60} -o gpurun_out/traffic_full -f \
   python tools/prof_run.py --ncrms 48602 --nstop 2 --calls 1 --spinup ${SPIN:-180} > gpurun_out/traffic.log 2>&1
tail -3 gpurun_out/traffic.log
