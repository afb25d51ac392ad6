#!/bin/bash
# BASELINE config 5: one configuration per process (fresh heap / NVLS binding each time)
N=${1:-8}
mkdir -p gpurun_out
: > gpurun_out/push_sweep_$N.jsonl
launch() { if [ "$N" = "1" ]; then timeout 120 python benchmarks/push_sweep.py "$@"; else timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) benchmarks/push_sweep.py "$@"; fi; }
for r in 1 2 4 8 16; do launch --ranks $r --budgets "" 2>/dev/null | grep push_sweep >> gpurun_out/push_sweep_$N.jsonl; done
for b in 0.01 0.05 0.25; do launch --ranks "" --budgets $b 2>/dev/null | grep push_sweep >> gpurun_out/push_sweep_$N.jsonl; done
cat gpurun_out/push_sweep_$N.jsonl
