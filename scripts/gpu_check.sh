#!/bin/bash
# One-GPU validation + evidence: gpu tests, smoke, headline bench, phase profile, ncu captures (run under gpurun).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench1.log 2>&1
timeout 300 python scripts/profile_step.py --channels-last --kernels --out gpurun_out/profile_bf16_cl_fusedbn.txt > /dev/null 2>gpurun_out/profile_err.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"skinny_gemm|bn_stats|bn_apply|bn_bwd|ext_finalize" -s 40 -c 12 -o gpurun_out/prof_tc_bn python scripts/ncu_ext.py ResNet18 > gpurun_out/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"skinny_gemm" -s 4 -c 4 -o gpurun_out/prof_gemm python scripts/ncu_ext.py VGG11 > gpurun_out/ncu3.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; grep -h '"metric"' gpurun_out/bench1.log | cut -c1-250; head -n 12 gpurun_out/profile_bf16_cl_fusedbn.txt; tail -n 2 gpurun_out/ncu2.log gpurun_out/ncu3.log
