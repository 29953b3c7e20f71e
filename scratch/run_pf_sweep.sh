#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for pf in 0 148 296 1184 2368; do EVP_K1_PF=$pf $B > gpurun_out/pf_$pf.log 2>&1; done
$B > gpurun_out/pf_default.log 2>&1
