#!/bin/bash
# Round-2 evidence in one call: ncu DRAM traffic at the full cfg2 batch, bench lines for cfg1 / cfg4 / cfg3 at their full nstop,
# and the pcols-sized sweep (32 / 256 / 2048 CRMs per call) of INTEGRATION.md section 4.
mkdir -p gpurun_out
bash tools/gpu_traffic.sh > gpurun_out/traffic_run.log 2>&1
python bench.py --config cfg1 2>gpurun_out/r02g_cfg1.err > gpurun_out/r02g_bench_cfg1_full.json
python bench.py --config cfg4 --steps 1 --warmup 3 2>gpurun_out/r02g_cfg4.err > gpurun_out/r02g_bench_cfg4_full.json
for n in 32 256 2048; do
  python bench.py --config cfg2 --ncrms $n --steps 3 --warmup 3 --no-cpu 2>>gpurun_out/r02g_sweep.err > gpurun_out/r02g_sweep_cfg2_${n}crms.json
  python bench.py --config cfg1 --ncrms $n --steps 3 --warmup 3 --no-cpu 2>>gpurun_out/r02g_sweep.err > gpurun_out/r02g_sweep_cfg1_${n}crms.json
done
python bench.py --config cfg3 --steps 1 --warmup 3 2>gpurun_out/r02g_cfg3.err > gpurun_out/r02g_bench_cfg3_full.json
for f in gpurun_out/r02g_*.json; do echo $f; cut -c1-260 $f; done
