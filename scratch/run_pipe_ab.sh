#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_pipe.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_pipe.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
EVP_PIPE3=0 $B > gpurun_out/pipe_off.log 2>&1
EVP_PIPE3=0 EVP_CHUNKS=4 $B > gpurun_out/pipe_off_c4.log 2>&1
$B > gpurun_out/pipe_c4.log 2>&1
EVP_CHUNKS=2 $B > gpurun_out/pipe_c2.log 2>&1
EVP_CHUNKS=8 $B > gpurun_out/pipe_c8.log 2>&1
EVP_CHUNKS=16 $B > gpurun_out/pipe_c16.log 2>&1
EVP_PIPE_PRIO=0 $B > gpurun_out/pipe_c4_prio0.log 2>&1
EVP_CHUNKS=8 EVP_PIPE_PRIO=0 $B > gpurun_out/pipe_c8_prio0.log 2>&1
$B --workload hcp > gpurun_out/pipe_hcp.log 2>&1
tail -n 3 gpurun_out/pytest_pipe.log
