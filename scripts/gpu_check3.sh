#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=200 -p no:cacheprovider -k "tcgen05" > gpurun_out/pytest_tc.log 2>&1
echo "exit $?" >> gpurun_out/pytest_tc.log
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider -k "not tcgen05_skinny" > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_tc.log; tail -n 40 gpurun_out/pytest_gpu.log
