#!/bin/bash
# scaling run on an 8-GPU box: N = 8, 4 (svd) + 8 (sgd dense NVLS path)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8.txt 2>&1
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/scale_${N}.log 2>&1
  echo "exit $?" >> gpurun_out/scale_${N}.log
  grep '"metric"' gpurun_out/scale_${N}.log | cut -c1-330
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 50 --warmup 5 --code sgd > gpurun_out/scale_8_sgd.log 2>&1
grep '"metric"' gpurun_out/scale_8_sgd.log | cut -c1-330
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 50 --warmup 5 --ps-mode dedicated > gpurun_out/scale_8_dedicated.log 2>&1
grep '"metric"' gpurun_out/scale_8_dedicated.log | cut -c1-330
tail -n 3 gpurun_out/scale_8.log | cut -c1-300
