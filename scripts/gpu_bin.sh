#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -k "golden or bucket_binning or capacity or sort_scan or c1_config or baseline_configs or vs_oracle" 2>&1 | tail -5
timeout 600 python scripts/ab_paths.py --config c3 --steps 5 --paths slab 2>&1 | tail -1 | cut -c1-700
for s in 1 2; do
timeout 600 python bench.py --config c5 --scale-mult $s --steps 5 --warmup 3 --no-cpu-baseline --no-graph 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('c5 x$s', round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items() if b>0.5})"
done
