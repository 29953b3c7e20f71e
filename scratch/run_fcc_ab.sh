#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_fcc.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_fcc.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/fcc_on.log 2>&1
EVP_K1_FCC=0 $B > gpurun_out/fcc_off.log 2>&1
$B --grid 512x512x512 --steps 5 > gpurun_out/fcc_512.log 2>&1
tail -n 3 gpurun_out/pytest_fcc.log
