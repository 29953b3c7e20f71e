#!/bin/bash
# round 2, GPU call 7 (one GPU): full -m gpu suite and the default bench with the 5x5 Newton (hydrostatic unknown eliminated)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/r02_c7_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02_c7_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_c7_bench.json 2> gpurun_out/r02_c7_bench.err
timeout 300 python bench.py --workload hcp --no-cpu-baseline --no-extras > gpurun_out/r02_c7_bench_hcp.json 2> gpurun_out/r02_c7_bench_hcp.err
for f in gpurun_out/r02_c7_bench*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']), [(k['name'], k['ms']) for k in d.get('kernels',[])], 'roof', d['roofline']['frac'], d['roofline'].get('achieved'), 'plastic', (d.get('plastic') or {}).get('ms_per_step'))
PY
done
