#!/bin/bash
# round 2, GPU call 5 (eight GPUs): 512^3 weak-scaling point: z kernel old/new, p2p wait mode, chunks, comm-stream priority, pencils, parity
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
B="--gpus 8 --no-cpu-baseline --no-extras --steps 10 --warmup 3"
run() { name=$1; shift; env "$@" timeout 300 $TR bench.py $B $EXTRA > gpurun_out/r02_c5_$name.json 2> gpurun_out/r02_c5_$name.err; }
EXTRA=""
timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_default.json 2> gpurun_out/r02_c5_default.err
run zk1 EVP_ZKERNEL=1
run waitread EVP_P2P_WAIT=read
run chunks2 EVP_CHUNKS=2
run prio1 EVP_COMM_PRIO=1
run priom1 EVP_COMM_PRIO=-1
run nccl EVP_TRANSPORT=nccl
EXTRA="--decomp pencil --py 2"; run pencil2x4 A=1
EXTRA="--decomp pencil --py 4"; run pencil4x2 A=1
EXTRA=""
MGPU_TRANSPORT=p2p MGPU_PY=1 MGPU_TMP=/tmp timeout 300 $TR tests/mgpu_check.py > gpurun_out/r02_c5_mgpu_p2p.log 2>&1; tail -2 gpurun_out/r02_c5_mgpu_p2p.log
MGPU_TRANSPORT=nccl MGPU_PY=2 MGPU_TMP=/tmp timeout 300 $TR tests/mgpu_check.py > gpurun_out/r02_c5_mgpu_pencil2x4.log 2>&1; tail -2 gpurun_out/r02_c5_mgpu_pencil2x4.log
for f in gpurun_out/r02_c5_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']), [(k['name'], k['ms']) for k in d.get('kernels',[])], 'exch', d.get('exchange_ms'), 'parity', (d.get('parity_check') or {}).get('ok'))
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
