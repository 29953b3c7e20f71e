#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chunked_pipeline.py -m gpu -x -q -k variants > gpurun_out/pytest_persist.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_persist.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for rep in 1 2; do
  $B > gpurun_out/ps0_fcc_$rep.log 2>&1
  EVP_K1_PERSIST=1 $B > gpurun_out/ps1_fcc_$rep.log 2>&1
  $B --workload hcp > gpurun_out/ps0_hcp_$rep.log 2>&1
  EVP_K1_PERSIST=1 $B --workload hcp > gpurun_out/ps1_hcp_$rep.log 2>&1
done
tail -n 2 gpurun_out/pytest_persist.log
