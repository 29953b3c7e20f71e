#!/bin/bash
# round 2, GPU call 10 (two GPUs): transposes by the copy engines (EVP_WAYBACK=dma): parity and timing
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
EVP_WAYBACK=dma MGPU_TRANSPORT=p2p MGPU_PY=1 MGPU_TMP=/tmp timeout 300 $TR tests/mgpu_check.py > gpurun_out/r02_c10_mgpu_dma.log 2>&1; tail -2 gpurun_out/r02_c10_mgpu_dma.log
B="--gpus 2 --no-cpu-baseline --no-extras --steps 10 --warmup 3"
run() { name=$1; shift; env "$@" timeout 200 $TR bench.py $B > gpurun_out/r02_c10_$name.json 2> gpurun_out/r02_c10_$name.err; }
run dma4 EVP_WAYBACK=dma
run dma2 EVP_WAYBACK=dma EVP_CHUNKS=2
run dma4_priohi EVP_WAYBACK=dma EVP_COMM_PRIO=-1
for f in gpurun_out/r02_c10_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']), [(k['name'], k['ms']) for k in d.get('kernels',[])], 'exch', d.get('exchange_ms'))
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
