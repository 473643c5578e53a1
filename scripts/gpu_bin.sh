#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -k "golden or bucket_binning or capacity or sort_scan or c1_config or baseline_configs" 2>&1 | tail -5
timeout 600 python scripts/ab_paths.py --config c3 --steps 5 --paths slab 2>&1 | tail -1 | cut -c1-700
