#!/bin/bash
# One gpurun call: parity of the backward formulations, A/B timing at c3 / c2, ncu of the grouped backward.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest rasterization parity (all backward modes)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "rasterization or full_size or equals_loop" 2>&1 | tail -15 | tee gpurun_out/pytest_raster.log
echo "== A/B c3"
timeout 600 python scripts/ab_blend_bwd.py --config c3 --steps 5 2>&1 | tail -6 | tee gpurun_out/ab_c3.log
echo "== A/B c2 (D=17) and c3 at D=5"
timeout 300 python scripts/ab_blend_bwd.py --config c2 --steps 5 2>&1 | tail -4 | tee gpurun_out/ab_c2.log
timeout 300 python scripts/ab_blend_bwd.py --config c3 --steps 3 --d0 4 2>&1 | tail -4 | tee gpurun_out/ab_c3_d5.log
if [ "$1" = "ncu" ]; then
  echo "== ncu full capture of the grouped backward"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_bwd -s 1 -c 1 \
      -o gpurun_out/prof_bwd_gp -f python scripts/ab_blend_bwd.py --config c3 --steps 1 --modes gp:0 > gpurun_out/ncu_gp.log 2>&1
  tail -2 gpurun_out/ncu_gp.log
  D4_BWD_GP_CFG=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_bwd -s 1 -c 1 \
      -o gpurun_out/prof_bwd_gp1 -f python scripts/ab_blend_bwd.py --config c3 --steps 1 --modes gp:1 > gpurun_out/ncu_gp1.log 2>&1
  tail -2 gpurun_out/ncu_gp1.log
fi
ls -la gpurun_out
