#!/bin/bash
# Parity: reference src/evaluate_pytorch.sh — poll the shared checkpoint directory.
python -m atomo_b200.distributed_evaluator --eval-batch-size=10000 --eval-freq=200 --model-dir=output/models/ \
  --dataset=Cifar10 --network=ResNet18 "$@"
