#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1
echo "exit $?" >> gpurun_out/pytest_kernels.log
timeout 300 python scripts/profile_step.py --kernels --out gpurun_out/profile_bf16.txt > /dev/null 2>gpurun_out/profile_err.txt
timeout 300 python scripts/profile_step.py --channels-last --kernels --out gpurun_out/profile_bf16_cl.txt > /dev/null 2>>gpurun_out/profile_err.txt
timeout 300 python scripts/profile_step.py --dtype fp32 --out gpurun_out/profile_fp32.txt > /dev/null 2>>gpurun_out/profile_err.txt
timeout 300 python bench.py --steps 30 --warmup 5 --channels-last > gpurun_out/bench_cl.log 2>&1
tail -n 3 gpurun_out/pytest_kernels.log; head -n 14 gpurun_out/profile_bf16.txt gpurun_out/profile_bf16_cl.txt gpurun_out/profile_fp32.txt
