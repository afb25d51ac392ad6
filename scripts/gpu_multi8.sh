#!/bin/bash
# 8-GPU validation: W=8 correctness tests + headline bench at N=8 (sharded / colocated) and N=4
mkdir -p gpurun_out
ATOMO_TEST_WORLD8=1 timeout 600 python -m pytest tests/test_gpu_v2.py -m gpu -q --timeout=300 -p no:cacheprovider -k "multi and (sharded or dense)" > gpurun_out/pytest_v2_w8.log 2>&1
echo "exit $?" >> gpurun_out/pytest_v2_w8.log; tail -n 6 gpurun_out/pytest_v2_w8.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
b() { name=$1; n=$2; shift; shift; timeout 300 $T --nproc-per-node $n --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 100 --warmup 5 "$@" > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "$name exit $? $(grep -ho '"value": [0-9.]*' gpurun_out/bench_$name.log | head -2 | tr '\n' ' ') $(grep -ho '"ms_per_step": [0-9.]*' gpurun_out/bench_$name.log | head -2 | tr '\n' ' ') $(grep -ho '"phase_us": {[^}]*}' gpurun_out/bench_$name.log) $(grep -ho '"nvls_multicast": [a-z]*' gpurun_out/bench_$name.log)"; }
b n8_sharded 8
b n8_colocated 8 --ps-mode colocated
b n4_sharded 4
b n8_sgd 8 --code sgd
tail -3 gpurun_out/bench_n8_sharded.err
