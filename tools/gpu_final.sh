#!/bin/bash
# End-of-round evidence in one gpurun call: GPU parity tests, smoke(), developed-state kernel times, one ncu --set full
# capture of every kernel of a developed sub-step (4736 CRMs), and a config-5 geometry bench line (one GPU's share).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
SPIN=180 bash tools/gpu_kt.sh > /dev/null 2>&1; head -4 gpurun_out/ktimes_dev.log
SPIN=180 bash tools/gpu_ncu.sh 'k_(p[0-9]|mpdata|precip_[ab]|stats|kurant)' 16 14 > /dev/null 2>&1
timeout 900 python bench.py --config cfg5 --ncrms 24301 --steps 1 --warmup 2 --no-cpu --e2e-steps 0 > gpurun_out/bench_cfg5.log 2>&1; tail -1 gpurun_out/bench_cfg5.log | cut -c1-600
