#!/bin/bash
# 2-GPU validation: multi-GPU tests of the shadow engine (incl. protocol fuzzing) + rank-budget sweep (config 5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v2.py -m gpu -q --timeout=400 -p no:cacheprovider -k "multi or protocol" > gpurun_out/pytest_v2_multi.log 2>&1
echo "exit $?" >> gpurun_out/pytest_v2_multi.log
tail -n 6 gpurun_out/pytest_v2_multi.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for r in 1 2 4 8 16; do
  timeout 200 $T --master-port $((29700 + r)) bench.py --gpus 2 --steps 60 --warmup 3 --svd-rank $r --no-fp32-line > gpurun_out/sweep2_r$r.log 2> gpurun_out/sweep2_r$r.err
  echo "r=$r exit $? $(grep -ho '"value": [0-9.]*' gpurun_out/sweep2_r$r.log | head -1) $(grep -ho '"ms_per_step": [0-9.]*' gpurun_out/sweep2_r$r.log | head -1) $(grep -ho '"phase_us": {[^}]*}' gpurun_out/sweep2_r$r.log)"
done
timeout 200 $T --master-port 29731 bench.py --gpus 2 --steps 60 --warmup 3 --code qsvd --no-fp32-line > gpurun_out/sweep2_qsvd.log 2>&1
echo "qsvd exit $? $(grep -ho '"ms_per_step": [0-9.]*' gpurun_out/sweep2_qsvd.log | head -1) $(grep -ho '"phase_us": {[^}]*}' gpurun_out/sweep2_qsvd.log)"
