#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench1.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --network ResNet50 --dataset ImageNet --batch-size 32 --svd-rank 8 > gpurun_out/cfg4_1.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --network ResNet50 --dataset ImageNet --batch-size 32 --svd-rank 8 --subspace off > gpurun_out/cfg4_1_nosub.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"skinny_gemm" -s 4 -c 4 -o gpurun_out/prof_gemm2 python scripts/ncu_ext.py VGG11 > gpurun_out/ncu3.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300; grep -h '"metric"' gpurun_out/bench1.log gpurun_out/cfg4_1.log gpurun_out/cfg4_1_nosub.log | cut -c1-230
