#!/bin/bash
# compute-sanitizer passes over the small parity tests + the c5 stress configuration once
mkdir -p gpurun_out
echo "== memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout 1400 \
   -k "golden or empty or contract or units or camera or densify or bucket or deform_against" 2>&1 | tail -6 | tee gpurun_out/memcheck.log
echo "== racecheck (blend + binning on the smallest fixtures)"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout 1400 \
   -k "golden and rgb or empty" 2>&1 | tail -6 | tee gpurun_out/racecheck.log
echo "== c5 stress (1M Gaussians, N=13, K=16)"
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1200 | tee gpurun_out/bench_c5.log
