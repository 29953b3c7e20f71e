#!/bin/bash
# round 2, GPU call 13 (eight GPUs): the default 512^3 bench line with the pull transport (parity check inside)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
tail -c 1500 gpurun_out/r02_bench_8gpu.json; tail -3 gpurun_out/r02_bench_8gpu.err
