#!/bin/bash
# 8-GPU run: c3 weak + e2e + strong scaling (band-major rows and 2-D units), then the c5 stress sweep
# (BASELINE configs[3] / configs[4])
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/scale8_c3.err | tail -1 > gpurun_out/scale8_c3.json
python -c "
import json; d=json.load(open('gpurun_out/scale8_c3.json')); s=d['strong']
print('c3 weak', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'strong', round(s['ms_per_blurry_frame'],3), round(s['speedup_vs_1gpu'],2))"
for s in 0.5 1 2; do
  timeout 400 $TR bench.py --gpus 8 --config c5 --scale-mult $s --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/c5_s$s.err | tail -1 > gpurun_out/c5_s$s.json
  python -c "
import json; d=json.load(open('gpurun_out/c5_s$s.json')); s=d['strong']
print('c5 x$s isects/frame', round(d['n_isects_per_frame']), 'weak', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'strong', round(s['ms_per_blurry_frame'],3), round(s['speedup_vs_1gpu'],2))"
done
