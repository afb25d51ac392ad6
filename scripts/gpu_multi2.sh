#!/bin/bash
# 2-GPU validation: multi-GPU tests of both engines + bench at N=2 in the three PS topologies
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v2.py -m gpu -q --timeout=400 -p no:cacheprovider -k "multi" > gpurun_out/pytest_v2_multi.log 2>&1
echo "exit $?" >> gpurun_out/pytest_v2_multi.log
tail -n 15 gpurun_out/pytest_v2_multi.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in sharded colocated dedicated; do
  timeout 300 $T --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 --ps-mode $mode > gpurun_out/bench2_$mode.log 2> gpurun_out/bench2_$mode.err
  echo "$mode exit $? $(grep -ho '"value": [0-9.]*, "unit": "images/s", "n_gpus": 2' gpurun_out/bench2_$mode.log | head -1) $(grep -ho '"ms_per_step": [0-9.]*' gpurun_out/bench2_$mode.log | head -2 | tr '\n' ' ') $(grep -ho '"phase_us": {[^}]*}' gpurun_out/bench2_$mode.log)"
done
timeout 300 $T --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 --engine fused > gpurun_out/bench2_fused.log 2> gpurun_out/bench2_fused.err
echo "fused exit $? $(grep -ho '"ms_per_step": [0-9.]*' gpurun_out/bench2_fused.log | head -2 | tr '\n' ' ')"
tail -3 gpurun_out/bench2_sharded.err
