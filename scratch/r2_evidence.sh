#!/bin/bash
# round 2 evidence (one GPU): default bench line, reference arm, HCP line, ncu launch list of the bench command, full captures of
# the two kernels furthest from their roofline (constitutive, fused z pass)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > gpurun_out/r02_gpu.txt
timeout 400 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
timeout 200 python bench.py --workload hcp --no-cpu-baseline > gpurun_out/r02_bench_hcp.json 2>> gpurun_out/r02_bench.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 125 -c 45 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_constitutive -s 14 -c 1 -f -o gpurun_out/prof_r02_constitutive $B > gpurun_out/ncu_c.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_zfused -s 14 -c 1 -f -o gpurun_out/prof_r02_zfused $B > gpurun_out/ncu_z.log 2>&1
tail -c 400 gpurun_out/r02_bench.json; echo; tail -3 gpurun_out/r02_launches.csv
