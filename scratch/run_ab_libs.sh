#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for rep in 1 2; do
for v in head new; do
  cp scratch/lib_$v.so lapx_b200/libevpfft_b200.so
  $B > gpurun_out/ab_${v}_fcc_$rep.log 2>&1
  $B --workload hcp > gpurun_out/ab_${v}_hcp_$rep.log 2>&1
done; done
