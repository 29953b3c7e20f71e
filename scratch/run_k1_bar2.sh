#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_chunked_pipeline.py -m gpu -x -q > gpurun_out/pytest_bar2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_bar2.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bar2_fcc.log 2>&1
$B --workload hcp > gpurun_out/bar2_hcp.log 2>&1
tail -n 2 gpurun_out/pytest_bar2.log
