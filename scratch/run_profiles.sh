#!/bin/bash
# round-1 (second session) evidence: bench lines, ncu launch list, full captures of the hot kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > gpurun_out/r01b_gpu.txt
python bench.py > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r01b_bench_reference.json 2>> gpurun_out/r01b_bench.err
python bench.py --workload hcp --no-cpu-baseline > gpurun_out/r01b_bench_hcp.json 2>> gpurun_out/r01b_bench.err
python bench.py --no-cpu-baseline --cufft > gpurun_out/r01b_bench_cufft.json 2>> gpurun_out/r01b_bench.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 125 -c 45 --csv --log-file gpurun_out/r01b_launches.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_constitutive -s 14 -c 1 -f -o gpurun_out/prof_r01b_constitutive $B > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_zfused -s 14 -c 1 -f -o gpurun_out/prof_r01b_zfused $B > gpurun_out/ncu_z.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_xfwd|k_ypass|k_xinv" -s 56 -c 4 -f -o gpurun_out/prof_r01b_xy $B > gpurun_out/ncu_xy.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01b_smoke.log 2>&1
tail -n 1 gpurun_out/r01b_smoke.log
