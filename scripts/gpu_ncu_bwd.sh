#!/bin/bash
# ncu --set full of the backward blend kernel under the given D4_BWD modes (args: e.g. gp:0 gp:1 shfl)
mkdir -p gpurun_out
for m in "$@"; do
  tag=$(echo $m | tr ':' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_bwd -s 1 -c 1 \
      -o gpurun_out/prof_bwd_$tag -f python scripts/ab_blend_bwd.py --config c3 --steps 1 --modes $m > gpurun_out/ncu_$tag.log 2>&1
  tail -1 gpurun_out/ncu_$tag.log
done
ls -la gpurun_out
