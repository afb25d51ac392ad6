#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=200 -p no:cacheprovider -k "fused_bn" > gpurun_out/pytest_bn.log 2>&1
echo "exit $?" >> gpurun_out/pytest_bn.log
timeout 200 python bench.py --steps 40 --warmup 5 --fused-bn on > gpurun_out/bench_bn_on.log 2>&1
timeout 200 python bench.py --steps 40 --warmup 5 --fused-bn off > gpurun_out/bench_bn_off.log 2>&1
tail -n 25 gpurun_out/pytest_bn.log | cut -c1-200; grep -h '"metric"' gpurun_out/bench_bn_on.log gpurun_out/bench_bn_off.log | cut -c1-260
grep -h "Error\|error" gpurun_out/bench_bn_on.log | tail -n 5
