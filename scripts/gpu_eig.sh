#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_eig.py > gpurun_out/bench_eig.txt 2>&1; cat gpurun_out/bench_eig.txt | tail -20
for sw in 1 2; do timeout 200 python bench.py --steps 150 --warmup 3 --max-sweeps $sw > gpurun_out/b_sw$sw.log 2>&1; echo "sweeps $sw: $(grep -ho '"ms_per_step": [0-9.]*' gpurun_out/b_sw$sw.log | head -2 | tr '\n' ' ') $(grep -ho '"final_loss": [0-9.]*' gpurun_out/b_sw$sw.log)"; done
