#!/bin/bash
# round 2, GPU call 3 (one GPU): phase offset between the compute groups of k_zfused3, ncu capture of k_zfused3
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
B="--grid 256x256x512 --no-cpu-baseline --no-extras --steps 10 --warmup 3"
for sk in 0 100 200 400 800; do EVP_Z3_SKEW=$sk timeout 300 python bench.py $B > gpurun_out/r02_c3_skew$sk.json 2>gpurun_out/r02_c3_skew$sk.err; done
for f in gpurun_out/r02_c3_skew*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('ms/step %.4f' % d['ms_per_step'], [(k['name'], k['ms']) for k in d.get('kernels',[])])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_zfused3 -s 14 -c 1 -f -o gpurun_out/prof_r02_zfused3 python bench.py --grid 256x256x512 --no-cpu-baseline --no-extras --steps 3 --warmup 3 > gpurun_out/ncu_z3.log 2>&1; tail -3 gpurun_out/ncu_z3.log
