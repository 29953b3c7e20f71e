#!/bin/bash
# A/B of the constitutive kernel variants on one B200 (run through gpurun)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
./scratch/dfma_bench > gpurun_out/dfma_bench.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_k1p.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_k1p.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
EVP_K1_LEGACY=1 $B > gpurun_out/k1_legacy.log 2>&1
for cfg in "3 12" "3 6" "4 12" "4 6" "4 4"; do set -- $cfg; mb=$1; g=$2;
  EVP_K1_MINB=$mb EVP_K1_G=$g $B > gpurun_out/k1p_mb${mb}_g${g}.log 2>&1
done
$B --workload hcp > gpurun_out/k1p_hcp.log 2>&1
EVP_K1_LEGACY=1 $B --workload hcp > gpurun_out/k1_legacy_hcp.log 2>&1
tail -n 3 gpurun_out/pytest_k1p.log
