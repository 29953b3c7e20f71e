#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_chord.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_chord.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/chord_fcc.log 2>&1
EVP_K1_MINB=3 $B > gpurun_out/chord_fcc_mb3.log 2>&1
$B --workload hcp > gpurun_out/chord_hcp.log 2>&1
EVP_K1_LEGACY=1 $B > gpurun_out/chord_legacy.log 2>&1
tail -n 3 gpurun_out/pytest_chord.log
