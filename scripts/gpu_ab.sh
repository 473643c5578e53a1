#!/bin/bash
# A/B of blend kernel variants on the GPU box: AB_MODES="shfl gp:0 D4_HIT_MASKS=0" [AB_D5=1] bash scripts/gpu_ab.sh [test]
# (modes: shfl | gp[:cfg] | ENV=VAL[,ENV=VAL...]; see scripts/ab_blend_bwd.py)
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "== A/B c3"
timeout 600 python scripts/ab_blend_bwd.py --config c3 --steps 5 --modes $AB_MODES 2>&1 | tail -8 | tee gpurun_out/ab_c3.log
if [ -n "$AB_D5" ]; then
timeout 600 python scripts/ab_blend_bwd.py --config c3 --steps 3 --d0 4 --modes $AB_MODES 2>&1 | tail -8 | tee gpurun_out/ab_c3_d5.log
fi
if [ "$1" = "test" ]; then
echo "== pytest rasterization parity (all backward modes)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "rasterization or full_size or equals_loop" 2>&1 | tail -15 | tee gpurun_out/pytest_raster.log
fi
