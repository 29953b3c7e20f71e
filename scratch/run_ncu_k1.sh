#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_constitutive -s 14 -c 1 -f -o gpurun_out/prof_k1p_v2 $B > gpurun_out/ncu_k1p.log 2>&1
tail -n 2 gpurun_out/ncu_k1p.log
