#!/bin/bash
# round 2, last GPU call (one GPU): the full -m gpu suite on the final tree, ncu capture of the 24-system (HCP) Newton kernel, 512^3 on one GPU
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 170 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/r02_final_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_final_pytest_gpu.log
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_constitutive -s 14 -c 1 -f -o gpurun_out/prof_r02_constitutive_hcp python bench.py --workload hcp --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_ch.log 2>&1; tail -1 gpurun_out/ncu_ch.log
timeout 110 python bench.py --grid 512x512x512 --no-cpu-baseline --no-extras --steps 5 --warmup 3 > gpurun_out/r02_bench_1gpu_512.json 2> gpurun_out/r02_bench_1gpu_512.err; tail -c 300 gpurun_out/r02_bench_1gpu_512.json
