#!/bin/bash
# One ncu --set full capture (with source counters) of one developed-state sub-step, 4736 CRMs of the cfg2 geometry.
# The profiler range starts after the untimed spin-up call (tools/prof_run.py --spinup -> cudaProfilerStart).
mkdir -p gpurun_out
REGEX=${1:-'k_(p[0-9]|mpdata|precip_[ab]|stats|kurant)'}
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$REGEX" -s ${3:-14} -c ${2:-16} \
   -o gpurun_out/prof_dev -f python tools/prof_run.py --ncrms 4736 --nstop 4 --calls 1 --spinup ${SPIN:-180} > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log
ls -la gpurun_out/
