#!/bin/bash
# Canonical launch (parity: reference src/run_pytorch.sh: mpirun -n 3, ResNet18/Cifar10, bs 128, lr 0.01,
# momentum 0, svd-rank 3).  Here: torchrun over the GPUs of this node, fused NVLink engine.
NGPU=${NGPU:-$(nvidia-smi -L | wc -l)}
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU} --master-addr 127.0.0.1 --master-port 29500 \
  -m atomo_b200.distributed_nn \
  --lr=0.01 --momentum=0.0 --network=ResNet18 --dataset=Cifar10 --batch-size=128 --test-batch-size=200 \
  --comm-type=Bcast --num-aggregate=0 --eval-freq=200 --epochs=10 --max-steps=1000000 --svd-rank=3 \
  --quantization-level=4 --bucket-size=512 --code=svd --enable-gpu=1 --backend=p2p --dtype=bf16 \
  --train-dir=output/models/ "$@"
