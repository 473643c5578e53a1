#!/bin/bash
# tensor-core slab kernels: parity of the variants, then A/B timing on c3
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -k "golden or vs_oracle or baseline_configs or c1_config or hit_masks" 2>&1 | tail -15 | tee gpurun_out/pytest_tc.log
timeout 600 python scripts/ab_paths.py --config c3 --steps 5 --paths ${1:-slab:2:0 slab:2:1} 2>&1 | tail -4 | tee gpurun_out/ab_tc.log
