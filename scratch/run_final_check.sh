#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_final.log
python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --no-cpu-baseline --workload hcp --steps 10 > gpurun_out/bench_final_hcp.json 2>> gpurun_out/bench_final.err
EVP_K1_FCC=0 python bench.py --no-cpu-baseline --workload hcp --steps 10 > gpurun_out/bench_final_hcp_runtime_tables.json 2>> gpurun_out/bench_final.err
tail -n 3 gpurun_out/pytest_final.log; tail -n 2 gpurun_out/smoke_final.log
