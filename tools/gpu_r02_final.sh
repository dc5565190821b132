#!/bin/bash
# End-of-round-2 evidence at HEAD in one call: smoke(), ncu --set full summaries (2-D developed state, 3-D), ncu launch list of the
# default bench command, compute-sanitizer memcheck + racecheck on the 2-D, 3-D and m2005 paths.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
bash tools/gpu_ncu.sh > gpurun_out/ncu2d_run.log 2>&1
bash tools/gpu_ncu3d.sh > gpurun_out/ncu3d_run.log 2>&1
bash tools/gpu_launchlist.sh > gpurun_out/launchlist_run.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/racecheck_case.py all > gpurun_out/memcheck_all.txt 2>&1; tail -2 gpurun_out/memcheck_all.txt
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/racecheck_case.py all > gpurun_out/racecheck_all.txt 2>&1; tail -2 gpurun_out/racecheck_all.txt
ls -la gpurun_out | tail -12
