#!/bin/bash
# One-GPU validation + evidence: gpu tests, smoke, headline bench (with the nested fp32 line), p2p launcher log lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench1.log 2>gpurun_out/bench1.err
timeout 300 python -m atomo_b200.distributed_nn --backend p2p --network ResNet18 --dataset Cifar10 --synthetic 1 --train-len 4096 --test-len 512 --batch-size 128 --code svd --svd-rank 3 --dtype bf16 --enable-gpu 1 --max-steps 60 --log-interval 20 --eval-freq 40 --eval-batches 2 --train-dir gpurun_out/ckpt/ --lr 0.05 --momentum 0.9 > gpurun_out/launcher_p2p.log 2>&1
echo "launcher exit $?" >> gpurun_out/launcher_p2p.log
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; grep -h '"metric"' gpurun_out/bench1.log | cut -c1-2600; tail -2 gpurun_out/bench1.err; grep -E "Worker:|Master:|Test set|exit|rror" gpurun_out/launcher_p2p.log | tail -12; ls gpurun_out/ckpt | head; rm -rf gpurun_out/ckpt
