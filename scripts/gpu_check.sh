#!/bin/bash
# One-GPU validation + evidence: gpu tests, smoke, headline bench (with the nested fp32 line), other configs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench1.log 2>gpurun_out/bench1.err
timeout 300 python bench.py --steps 60 --warmup 3 --network VGG11 --code qsgd --no-fp32-line > gpurun_out/bench_cfg3_vgg11_qsgd_1.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 3 --network ResNet50 --dataset ImageNet --batch-size 32 --svd-rank 8 --no-fp32-line > gpurun_out/bench_cfg4_resnet50_1.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 3 --network ResNet50 --dataset ImageNet --batch-size 32 --code sgd --no-fp32-line > gpurun_out/bench_cfg4_resnet50_sgd_1.log 2>&1
timeout 300 python bench.py --steps 15 --warmup 3 --impl nccl-baseline --dtype bf16 > gpurun_out/bench_ncclbase_bf16_1.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; for f in bench1 bench_cfg3_vgg11_qsgd_1 bench_cfg4_resnet50_1 bench_cfg4_resnet50_sgd_1 bench_ncclbase_bf16_1; do echo "$f: $(grep -ho '"value": [0-9.]*' gpurun_out/$f.log | head -3 | tr '\n' ' ') $(grep -ho '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -3 | tr '\n' ' ')"; tail -1 gpurun_out/$f.log | cut -c1-200 | grep -v metric; done
