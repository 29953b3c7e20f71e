#!/bin/bash
# round 2, GPU call 9 (two GPUs): pull mode: peer-store wait mode and transpose-stream priority
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
B="--gpus 2 --no-cpu-baseline --no-extras --steps 10 --warmup 3"
run() { name=$1; shift; env "$@" timeout 200 $TR bench.py $B > gpurun_out/r02_c9_$name.json 2> gpurun_out/r02_c9_$name.err; }
run base A=1
run waitread EVP_P2P_WAIT=read
run priohi EVP_COMM_PRIO=-1
run priolo EVP_COMM_PRIO=1
run waitread_priohi EVP_P2P_WAIT=read EVP_COMM_PRIO=-1
for f in gpurun_out/r02_c9_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']), [(k['name'], k['ms']) for k in d.get('kernels',[])])
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
