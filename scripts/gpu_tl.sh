#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 150 --warmup 3 --groups 5"
r() { name=$1; shift; timeout 200 $B "$@" > gpurun_out/p_$name.log 2>&1; echo "$name: $(grep -ho '"ms_per_step": [0-9.]*' gpurun_out/p_$name.log | head -2 | tr '\n' ' ')"; }
r side_hi
r equal --side-priority 0
r main_hi --main-priority -1 --side-priority 0
r grid148 --ps-grid 148
r grid148_mainhi --ps-grid 148 --main-priority -1 --side-priority 0
r grid74 --ps-grid 74
r g4 --groups 4
r g4_equal --groups 4 --side-priority 0
