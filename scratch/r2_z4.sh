#!/bin/bash
# round 2 (one GPU, last GPU seconds): k_zfused4 (nz = 256) against the one-shot kernel and k_zfused2
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 60 python -m pytest tests/test_chunked_pipeline.py tests/test_gpu_parity.py -m gpu -q -k "persistent or full_size" --maxfail=3 > gpurun_out/r02_z4_pytest.log 2>&1; tail -3 gpurun_out/r02_z4_pytest.log
B="--no-cpu-baseline --no-extras --steps 10 --warmup 3"
EVP_Z4=1 timeout 45 python bench.py $B > gpurun_out/r02_z4_on.json 2> gpurun_out/r02_z4_on.err
EVP_Z4=0 timeout 45 python bench.py $B > gpurun_out/r02_z4_off.json 2> gpurun_out/r02_z4_off.err
for f in gpurun_out/r02_z4_on.json gpurun_out/r02_z4_off.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.4f' % d['ms_per_step'], [(k['name'], k['ms'], k['frac_hbm']) for k in d['kernels']])
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
done
