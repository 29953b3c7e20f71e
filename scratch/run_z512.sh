#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_z512.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_z512.log
B="python bench.py --warmup 3 --no-cpu-baseline"
$B --grid 512x512x512 --steps 5 > gpurun_out/z512_nb2.log 2>&1
EVP_ZNB=1 $B --grid 512x512x512 --steps 5 > gpurun_out/z512_nb1.log 2>&1
EVP_ZKERNEL=1 $B --grid 512x512x512 --steps 5 > gpurun_out/z512_oneshot.log 2>&1
$B --steps 10 > gpurun_out/z256_now.log 2>&1
tail -n 3 gpurun_out/pytest_z512.log
