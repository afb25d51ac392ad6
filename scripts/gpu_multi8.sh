#!/bin/bash
# 8-GPU run: headline bench exactly as the driver launches it (incl. the nested fp32 child), configs 3 and 4
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
b() { name=$1; n=$2; shift; shift; timeout 420 $T --nproc-per-node $n --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "$name exit $? $(grep -ho '"value": [0-9.]*' gpurun_out/bench_$name.log | head -3 | tr '\n' ' ') $(grep -ho '"ms_per_step": [0-9.]*' gpurun_out/bench_$name.log | head -3 | tr '\n' ' ') $(grep -ho '"phase_us": {[^}]*}' gpurun_out/bench_$name.log)"; }
b n8_final 8 --steps 100 --warmup 5
b n8_cfg4_resnet50 8 --steps 40 --warmup 3 --network ResNet50 --dataset ImageNet --batch-size 32 --svd-rank 8 --no-fp32-line
b n8_cfg3_vgg11_qsgd 8 --steps 60 --warmup 3 --network VGG11 --code qsgd --no-fp32-line
tail -3 gpurun_out/bench_n8_final.err
