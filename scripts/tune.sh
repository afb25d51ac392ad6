#!/bin/bash
# LR grid search (parity: reference src/tune.sh: 7 learning rates x 100 steps, then tiny_tuning_parser).
NGPU=${NGPU:-$(nvidia-smi -L | wc -l)}
NW=${NGPU}
mkdir -p tune
for lr in 0.0078125 0.015625 0.03125 0.0625 0.125 0.25 0.5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU} --master-addr 127.0.0.1 --master-port 29500 \
    -m atomo_b200.distributed_nn --lr=${lr} --momentum=0.9 --network=ResNet18 --dataset=Cifar10 --batch-size=8 \
    --eval-freq=100000 --max-steps=100 --log-interval=100 --svd-rank=3 --code=svd --enable-gpu=1 --backend=p2p --dtype=bf16 \
    --train-dir=output/models/ > tune/raw_${lr} 2>&1
  grep "Step: 100," tune/raw_${lr} > tune/${lr}
  python -m atomo_b200.tiny_tuning_parser --tuning-dir=tune/ --tuning-lr=${lr} --num-workers=${NW}
done
