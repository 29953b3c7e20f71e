#!/bin/bash
# round 2, GPU call 8 (two GPUs): pull way-back with the slab-wide receive buffer: parity (p2p + nccl), chunks 8 / 4 / 2
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
MGPU_TRANSPORT=p2p MGPU_PY=1 MGPU_TMP=/tmp timeout 300 $TR tests/mgpu_check.py > gpurun_out/r02_c8_mgpu_pull.log 2>&1; tail -2 gpurun_out/r02_c8_mgpu_pull.log
B="--gpus 2 --no-cpu-baseline --no-extras --steps 10 --warmup 3"
timeout 300 $TR bench.py --gpus 2 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r02_c8_chunks8.json 2> gpurun_out/r02_c8_chunks8.err
EVP_CHUNKS=4 timeout 300 $TR bench.py $B > gpurun_out/r02_c8_chunks4.json 2> gpurun_out/r02_c8_chunks4.err
EVP_CHUNKS=2 timeout 300 $TR bench.py $B > gpurun_out/r02_c8_chunks2.json 2> gpurun_out/r02_c8_chunks2.err
for f in gpurun_out/r02_c8_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']), [(k['name'], k['ms']) for k in d.get('kernels',[])], 'exch', d.get('exchange_ms'), 'parity', (d.get('parity_check') or {}).get('ok'), 'hcp', (d.get('config4_hcp') or {}).get('ms_per_step'))
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
