#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_zfused -s 14 -c 1 -f -o gpurun_out/prof_z2_v1 $B > gpurun_out/ncu_z2.log 2>&1
tail -n 2 gpurun_out/ncu_z2.log
