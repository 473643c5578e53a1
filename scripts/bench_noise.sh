#!/bin/bash
# run-to-run spread of the bench legs: N repeats of the default invocation
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e'].get('cuda_mallocs_in_timed_region'), 'graph', round(d['graph']['ms_per_step'],3) if d.get('graph') else None, d['clocks'])" "$1"; }
for i in $(seq 1 ${1:-6}); do
  python bench.py --steps ${2:-20} --warmup 3 --no-cpu-baseline 2>&1 | show "run$i"
done
