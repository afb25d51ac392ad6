#!/bin/bash
# One-shot GPU validation used under gpurun: probe, tests, smoke, short bench.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/nvidia_smi.txt 2>&1
python -c "
import torch, atomo_b200._C as C
print('cuda', torch.cuda.is_available(), torch.cuda.device_count(), torch.cuda.get_device_name(0))
print('mc_supported', C.heap_multicast_supported(0), 'posix_fd', C.heap_posix_fd_supported(0))
" > gpurun_out/probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench1.log 2>&1
echo "bench exit $?" >> gpurun_out/bench1.log
tail -5 gpurun_out/probe.txt gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench1.log
