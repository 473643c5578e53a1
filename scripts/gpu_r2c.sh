#!/bin/bash
# quick: correctness of the raster paths on the small scenes + c3 A/B of the slab path
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "golden_fixtures or c1_config or hit_masks or vs_oracle" 2>&1 | tail -8 | tee gpurun_out/pytest_raster.log
timeout 300 python scripts/ab_paths.py --config c3 --steps 5 --paths ${AB_PATHS:-slab} 2>&1 | tail -4 | tee gpurun_out/ab_c3.log
timeout 300 python scripts/ab_paths.py --config c3 --steps 3 --d0 4 --paths ${AB_PATHS:-slab} 2>&1 | tail -4 | tee gpurun_out/ab_c3_d5.log
