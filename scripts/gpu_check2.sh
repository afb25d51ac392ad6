#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/profile_step.py --channels-last --out gpurun_out/profile_bf16_cl.txt > /dev/null 2>gpurun_out/profile_err.txt
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench1.log 2>&1
# ncu: full capture of our three hottest kernels (one GPU, few launches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ps_update_kernel|project_push_kernel|gram_kernel|eig_sample_kernel" -s 8 -c 8 -o gpurun_out/prof_kernels python bench.py --steps 3 --warmup 3 --no-graph > gpurun_out/ncu.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; head -n 12 gpurun_out/profile_bf16_cl.txt; grep '"metric"' gpurun_out/bench1.log | cut -c1-250; tail -n 3 gpurun_out/ncu.log
