#!/bin/bash
mkdir -p gpurun_out
EVP_ZNB=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_chunked_pipeline.py -m gpu -x -q > gpurun_out/pytest_z.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_z.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/z_nb1.log 2>&1
EVP_ZNB=2 $B > gpurun_out/z_nb2.log 2>&1
EVP_ZNB=2 $B --grid 128x128x128 > gpurun_out/z_nb2_128.log 2>&1
tail -n 3 gpurun_out/pytest_z.log
