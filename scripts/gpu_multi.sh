#!/bin/bash
# multi-GPU validation: heap (vmm + NVLS multicast), replicas identical, N-GPU bench
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -q --timeout=400 -p no:cacheprovider -k "two_gpu" -s > gpurun_out/pytest_multi_${N}.log 2>&1
echo "exit $?" >> gpurun_out/pytest_multi_${N}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/bench_${N}.log 2>&1
echo "exit $?" >> gpurun_out/bench_${N}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 40 --warmup 5 --code sgd > gpurun_out/bench_${N}_sgd.log 2>&1
tail -n 12 gpurun_out/pytest_multi_${N}.log; tail -n 3 gpurun_out/bench_${N}.log | cut -c1-1500; tail -n 2 gpurun_out/bench_${N}_sgd.log | cut -c1-600
