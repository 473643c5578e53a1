#!/bin/bash
# capacity-mode tests + bench + ncu of the binning / deform kernels
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "capacity or empty or frame_renderer" 2>&1 | tail -5 | tee gpurun_out/pytest_cap.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_sort|bucket_emit|tile_count|deform_fg_bwd|combine" -s 12 -c 6 \
    -o gpurun_out/r02_bin -f python scripts/ab_paths.py --config c3 --steps 1 --paths slab > gpurun_out/ncu_bin.log 2>&1
tail -2 gpurun_out/ncu_bin.log
