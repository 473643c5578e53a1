#!/bin/bash
# round 2: slab-path parity + A/B timing + ncu capture of the slab blend kernels
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "== pytest raster parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "rasterization or hit_masks or empty or caller or equals_loop or bucket" 2>&1 | tail -25 | tee gpurun_out/pytest_raster.log
echo "== A/B c3"
timeout 300 python scripts/ab_paths.py --config c3 --steps 5 2>&1 | tail -6 | tee gpurun_out/ab_c3.log
timeout 300 python scripts/ab_paths.py --config c3 --steps 3 --d0 4 2>&1 | tail -6 | tee gpurun_out/ab_c3_d5.log
if [ "$1" = "ncu" ]; then
echo "== ncu full capture of the slab blend kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 6 -c 2 \
    -o gpurun_out/r02_blend -f python scripts/ab_paths.py --config c3 --steps 1 --paths slab > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
