#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_mgpu2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_new.log 2>&1
EVP_TRANSPORT=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_new_nccl.log 2>&1
tail -n 3 gpurun_out/pytest_mgpu2.log
tail -n 1 gpurun_out/bench_2gpu_new.log | cut -c1-300
