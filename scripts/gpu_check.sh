#!/bin/bash
# One gpurun call: parity tests, smoke, short bench, ncu launch list (+ optional full capture).
# usage: scripts/gpu_check.sh [quick|full]
mode=${1:-full}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
if [ "$mode" = "full" ]; then
  echo "== ncu launch list (every kernel of 3 warm-up + 2 timed + 4 e2e steps)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-fwd-only > gpurun_out/ncu_bench.log 2>&1
  tail -2 gpurun_out/ncu_bench.log
  echo "== ncu full capture of the blend kernels"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 6 -c 2 \
      -o gpurun_out/r02_blend_final -f python scripts/ab_paths.py --config c3 --steps 1 --paths slab > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
  ls -la gpurun_out | tail -5
fi
