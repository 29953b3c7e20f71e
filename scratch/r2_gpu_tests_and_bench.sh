#!/bin/bash
# round 2, GPU call 1 (one GPU): smoke, the full -m gpu suite, the default bench, the reference arm, the HCP workload
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 --durations=15 > gpurun_out/r02_pytest_gpu.log 2>&1; tail -40 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 3000 gpurun_out/r02_bench.json; tail -5 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference.json 2>&1; tail -c 600 gpurun_out/r02_bench_reference.json
timeout 600 python bench.py --workload hcp --no-cpu-baseline > gpurun_out/r02_bench_hcp.json 2> gpurun_out/r02_bench_hcp.err; tail -c 1500 gpurun_out/r02_bench_hcp.json
