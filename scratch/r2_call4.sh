#!/bin/bash
# round 2, GPU call 4 (one GPU): in-place k_zfused3 (row-swap layout): memcheck on a tiny case, parity, timing, ncu
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
cat > /tmp/tiny512.py <<'PY'
import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from common import make_polycrystal, rel_err
from lapx_b200 import api
lib = api.load_product()
outs = []
for flags in (0, 4):
    s, ids, grot = make_polycrystal(lib, lib, (16, 16, 512), 12, seed=5)
    s.set_profiling(flags)
    s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=10**6, tol_newton=1e-9, newton_itmax=100)
    s.set_loading(api.Loading.uniaxial_tension(1.0))
    s.begin_increment(2e-4)
    for it in range(3):
        r = s.equilibrium_iter()
    outs.append(s.get_field(api.FIELD_STRESS))
print("tiny512 rel diff new vs one-shot:", rel_err(outs[0], outs[1]))
PY
timeout 120 python /tmp/tiny512.py > gpurun_out/r02_c4_tiny.log 2>&1; tail -3 gpurun_out/r02_c4_tiny.log
if ! grep -q "rel diff" gpurun_out/r02_c4_tiny.log; then
  timeout 300 compute-sanitizer --tool memcheck python /tmp/tiny512.py > gpurun_out/r02_c4_memcheck.log 2>&1; grep -m 12 -E "Invalid|Error|at 0x|k_zfused|misaligned" gpurun_out/r02_c4_memcheck.log
  exit 0
fi
timeout 600 python -m pytest tests/test_chunked_pipeline.py tests/test_gpu_parity.py -m gpu -q -k "persistent or fixed_iterations" --maxfail=3 > gpurun_out/r02_c4_pytest_z.log 2>&1; tail -6 gpurun_out/r02_c4_pytest_z.log
B="--grid 256x256x512 --no-cpu-baseline --no-extras --steps 10 --warmup 3"
timeout 300 python bench.py $B > gpurun_out/r02_c4_nz512.json 2>gpurun_out/r02_c4_nz512.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c4_nz512.json').read().strip().splitlines()[-1])
print('ms/step %.4f' % d['ms_per_step'], [(k['name'], k['ms'], k['frac_hbm']) for k in d.get('kernels',[])])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_zfused3 -s 14 -c 1 -f -o gpurun_out/prof_r02_zfused3b python bench.py --grid 256x256x512 --no-cpu-baseline --no-extras --steps 3 --warmup 3 > gpurun_out/ncu_z3b.log 2>&1; tail -2 gpurun_out/ncu_z3b.log
