#!/bin/bash
# ncu --set full capture of one sub-step of the 3-D CRM (cfg4 geometry, 592 CRMs = 4 blocks per SM) after a short spin-up.
mkdir -p gpurun_out
REGEX=${1:-'k3?_(p[0-9]|mpdata|precip_[ab]|stats|kurant)'}
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$REGEX" -s ${3:-0} -c ${2:-16} \
   -o gpurun_out/prof_3d -f python tools/prof_run.py --config cfg4 --ncrms ${NCRMS:-592} --nstop 2 --calls 1 --spinup ${SPIN:-40} > gpurun_out/prof3d.log 2>&1
tail -3 gpurun_out/prof3d.log
ls -la gpurun_out/ | tail -5
