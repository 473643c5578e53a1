#!/bin/bash
# multi-GPU check: parity of the strong partitions against one GPU, then the bench (weak + strong legs) on N ranks
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/strong_check.py --config c2 2>&1 | grep -E '"mode"|Error|error' | cut -c1-400 | tee gpurun_out/strong_check_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps ${2:-10} --warmup 3 --no-cpu-baseline 2> gpurun_out/scale$N.err | tail -1 > gpurun_out/scale$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/scale$N.json"))
s = d["strong"]
print("weak", d["value"], d["ms_per_step"], d["e2e"])
print("strong rows", {k: s[k] for k in ["ms_per_blurry_frame", "speedup_vs_1gpu", "collective_ms", "collective_ms_by_tag"]})
a = s["alt_2d_units"]
print("strong 2d", {k: a[k] for k in ["ms_per_blurry_frame", "speedup_vs_1gpu", "collective_ms"]})
print({k: round(v, 3) for k, v in s["kernel_ms_per_step"].items()})
PY
tail -5 gpurun_out/scale$N.err
