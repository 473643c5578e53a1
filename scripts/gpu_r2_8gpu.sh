#!/bin/bash
# 8-GPU run: c3 weak + e2e + strong (bands) scaling, then the c5 stress sweep (BASELINE configs[3] / configs[4])
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/scale8_c3.err | tail -1 > gpurun_out/scale8_c3.json
tail -c 600 gpurun_out/scale8_c3.json; echo
for s in 0.5 1 2; do
  timeout 400 $TR bench.py --gpus 8 --config c5 --scale-mult $s --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/c5_s$s.err | tail -1 > gpurun_out/c5_s$s.json
  tail -c 300 gpurun_out/c5_s$s.json; echo
done
