#!/bin/bash
# round 2, GPU call 11 (four GPUs): pull mode + pencils (2 x 2) against the oracle, the default 4-GPU bench line (with parity check and the config-4 HCP leg)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544"
MGPU_TRANSPORT=p2p MGPU_PY=1 MGPU_TMP=/tmp timeout 200 $TR tests/mgpu_check.py > gpurun_out/r02_c11_mgpu_pull4.log 2>&1; tail -1 gpurun_out/r02_c11_mgpu_pull4.log
MGPU_TRANSPORT=nccl MGPU_PY=2 MGPU_TMP=/tmp timeout 200 $TR tests/mgpu_check.py > gpurun_out/r02_c11_mgpu_pencil2x2.log 2>&1; tail -1 gpurun_out/r02_c11_mgpu_pencil2x2.log
timeout 300 $TR bench.py --gpus 4 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r02_c11_4gpu.json 2> gpurun_out/r02_c11_4gpu.err
timeout 200 $TR bench.py --gpus 4 --no-cpu-baseline --no-extras --steps 10 --warmup 3 --decomp pencil > gpurun_out/r02_c11_4gpu_pencil.json 2> gpurun_out/r02_c11_4gpu_pencil.err
for f in gpurun_out/r02_c11_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']), [(k['name'], k['ms']) for k in d.get('kernels',[])], 'exch', d.get('exchange_ms'), 'parity', (d.get('parity_check') or {}).get('ok'), 'hcp', (d.get('config4_hcp') or {}).get('ms_per_step'))
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
