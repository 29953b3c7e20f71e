#!/bin/bash
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool memcheck python scratch/sanitize_small.py > gpurun_out/r01b_sanitizer_memcheck.log 2>&1
timeout 200 compute-sanitizer --tool racecheck python scratch/sanitize_small.py > gpurun_out/r01b_sanitizer_racecheck.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_final.log
python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -n 2 gpurun_out/r01b_sanitizer_memcheck.log; tail -n 2 gpurun_out/r01b_sanitizer_racecheck.log; tail -n 2 gpurun_out/pytest_final.log
