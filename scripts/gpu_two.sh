#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q --timeout=300 -p no:cacheprovider -k "two_gpu" > gpurun_out/pytest_multi_2.log 2>&1
echo "exit $?" >> gpurun_out/pytest_multi_2.log
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29611 bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/b2_bench_on.log 2>&1
timeout 300 $R --master-port 29612 bench.py --gpus 2 --steps 60 --warmup 5 --no-cudnn-benchmark > gpurun_out/b2_bench_off.log 2>&1
timeout 200 bash scripts/push_sweep.sh 2 > gpurun_out/sweep2_stdout.log 2>&1
tail -n 3 gpurun_out/pytest_multi_2.log; for f in b2_bench_on b2_bench_off; do grep -h '"metric"' gpurun_out/$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phase_us'], d['e2e']['value'])"; done; cat gpurun_out/push_sweep_2.jsonl | cut -c1-330
