#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err
python bench.py --workload hcp --no-cpu-baseline > gpurun_out/r01b_bench_hcp.json 2>> gpurun_out/r01b_bench.err
python bench.py --no-cpu-baseline --cufft > gpurun_out/r01b_bench_cufft.json 2>> gpurun_out/r01b_bench.err
tail -c 300 gpurun_out/r01b_bench.err
