#!/bin/bash
# ncu full capture of one steady sub-step in the DEVELOPED state (180-step spin-up, profiler started after it), 4736 CRMs
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:'k_(p[0-9]|mpdata|precip_[ab]|stats|kurant)' -s 15 -c 16 \
   -o gpurun_out/prof_cur -f python tools/prof_run.py --ncrms 4736 --spinup 180 --nstop 6 --calls 1 > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log
