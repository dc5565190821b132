#!/bin/bash
# Final bench lines of round 2 at HEAD: full GPU test suite, default cfg2 (both arms), cfg4 and cfg3 at their full nstop.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -2
python bench.py 2>gpurun_out/r02i_cfg2.err > gpurun_out/r02i_bench_cfg2_full.json
python bench.py --impl reference 2>>gpurun_out/r02i_cfg2.err > gpurun_out/r02i_bench_cfg2_reference_arm.json
python bench.py --config cfg4 --steps 1 --warmup 3 2>gpurun_out/r02i_cfg4.err > gpurun_out/r02i_bench_cfg4_full.json
python bench.py --config cfg3 --steps 1 --warmup 3 2>gpurun_out/r02i_cfg3.err > gpurun_out/r02i_bench_cfg3_full.json
for f in gpurun_out/r02i_*.json; do echo $f; cut -c1-200 $f; done
