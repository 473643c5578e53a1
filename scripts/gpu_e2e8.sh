#!/bin/bash
# 8 GPUs, c3: end-to-end leg with and without the allocator's split limit
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], 'weak', round(d['ms_per_step'],3), 'e2e', d['e2e'])" "$1"; }
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-strong 2>/dev/null | show "split-limit(default), no strong leg"
PYTORCH_CUDA_ALLOC_CONF=garbage_collection_threshold:0.99 timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-strong 2>/dev/null | show "no split limit"
