#!/bin/bash
# End-of-round-2 evidence at HEAD in one call: smoke(), ncu --set full summaries (2-D developed state, 3-D) reduced to markdown on the
# box (the .ncu-rep files exceed the 64 MiB that gpurun copies back), ncu launch list of the default bench command.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
bash tools/gpu_ncu.sh > gpurun_out/ncu2d_run.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_dev.ncu-rep > gpurun_out/r02h_developed_state_ncu_full.md 2>/dev/null
for k in k_mpdata k_p8_micro k_precip_b; do python tools/sass_lines.py gpurun_out/prof_dev.ncu-rep $k ${k}ILi64 25 > gpurun_out/r02h_lines_$k.txt 2>&1; done
rm -f gpurun_out/prof_dev.ncu-rep
bash tools/gpu_ncu3d.sh > gpurun_out/ncu3d_run.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_3d.ncu-rep > gpurun_out/r02h_3d_cfg4_ncu_full.md 2>/dev/null
rm -f gpurun_out/prof_3d.ncu-rep
bash tools/gpu_launchlist.sh > gpurun_out/launchlist_run.log 2>&1
ls -la gpurun_out | tail -14
