#!/bin/bash
# Evidence run (1 GPU): ncu --set full of the v2 kernels + BN kernels inside the real step, compute-sanitizer logs
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"v2_encode|v2_project|v2_ps_kernel" -s 65 -c 15 -o gpurun_out/prof_v2 python scripts/profile_shadow.py --overlap --groups 5 > gpurun_out/ncu_v2.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"bn_stats|bn_apply|bn_bwd" -s 160 -c 8 -o gpurun_out/prof_bn python scripts/profile_shadow.py --groups 5 > gpurun_out/ncu_bn.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_v2.py -m gpu -q -x -p no:cacheprovider -k "gram or topk or ps_matches or qsvd or num_aggregate" > gpurun_out/sanitizer_v2_$tool.log 2>&1
  echo "v2 $tool exit $? : $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed' gpurun_out/sanitizer_v2_$tool.log | tr '\n' ' ')"
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "gram or full_rank or ps_update or qsgd or entrywise or bn" > gpurun_out/sanitizer_v1_memcheck.log 2>&1
echo "v1 memcheck exit $? : $(grep -h 'ERROR SUMMARY\|passed\|failed' gpurun_out/sanitizer_v1_memcheck.log | tr '\n' ' ')"
tail -2 gpurun_out/ncu_v2.log gpurun_out/ncu_bn.log
