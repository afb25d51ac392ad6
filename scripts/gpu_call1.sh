#!/bin/bash
# round-2 call 1: gpu tests (incl. the moved next-round tests), bench N=1 with the fixed timing, bf16-leaf probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench1.log 2>gpurun_out/bench1.err
timeout 600 python scripts/probe_bf16_leaf.py > gpurun_out/probe.log 2>gpurun_out/probe.err
tail -n 6 gpurun_out/pytest_gpu.log; grep -h '"metric"' gpurun_out/bench1.log | cut -c1-1500; tail -n 3 gpurun_out/bench1.err; cat gpurun_out/probe.log; tail -n 5 gpurun_out/probe.err
