#!/bin/bash
# BASELINE.json configs 3, 4, 5 (+ NCCL baseline) on N GPUs
N=${1:-1}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "1" ]; then RUN="python"; PORTARG=""; fi
run() { name=$1; shift; if [ "$N" = "1" ]; then timeout 500 python "$@" > gpurun_out/${name}_$N.log 2>&1; else timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@" > gpurun_out/${name}_$N.log 2>&1; fi; echo "exit $?" >> gpurun_out/${name}_$N.log; grep -h '"metric"\|push_sweep\|Error\|error' gpurun_out/${name}_$N.log | cut -c1-420 | tail -n 12; }
run cfg3_vgg11_qsgd bench.py --gpus $N --steps 30 --warmup 5 --network VGG11 --code qsgd --quantization-level 4
run cfg4_resnet50_imagenet bench.py --gpus $N --steps 20 --warmup 5 --network ResNet50 --dataset ImageNet --batch-size 32 --svd-rank 8
timeout 900 bash scripts/push_sweep.sh $N | cut -c1-330   # one configuration per process (fresh NVLS binding each)
run cfg2_subspace bench.py --gpus $N --steps 40 --warmup 5
if [ "$N" != "1" ]; then run baseline bench.py --gpus $N --steps 8 --warmup 3 --impl nccl-baseline; fi
