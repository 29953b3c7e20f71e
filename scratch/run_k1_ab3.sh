#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_k1p.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_k1p.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
EVP_K1_LEGACY=1 $B > gpurun_out/k1_legacy.log 2>&1
$B > gpurun_out/k1p_default.log 2>&1
EVP_K1_MINB=3 $B > gpurun_out/k1p_mb3.log 2>&1
$B --workload hcp > gpurun_out/k1p_hcp.log 2>&1
EVP_K1_G=24 $B --workload hcp > gpurun_out/k1p_hcp_g24.log 2>&1
tail -n 3 gpurun_out/pytest_k1p.log
