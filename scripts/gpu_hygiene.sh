#!/bin/bash
# compute-sanitizer passes over the small parity tests (all blend formulations incl. the tensor-core kernels, the merge
# path of the tile sort, the cost-volume kernels)
mkdir -p gpurun_out
echo "== memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout 1400 \
   -k "golden or empty or contract or units or camera or densify or bucket or deform_against or correlation or hit_masks" 2>&1 | tail -6 | tee gpurun_out/memcheck.log
echo "== racecheck (blend + binning on the smallest fixtures, every formulation)"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout 1400 \
   -k "golden and (rgb or ed17) or empty" 2>&1 | tail -6 | tee gpurun_out/racecheck.log
echo "== synccheck"
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout 1400 \
   -k "golden and ed17 or bucket" 2>&1 | tail -6 | tee gpurun_out/synccheck.log
