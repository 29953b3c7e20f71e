#!/bin/bash
# round 2, GPU call 2 (two GPUs): nz = 512 z kernel, N ranks vs oracle (p2p / nccl / pencil), 2-GPU bench A/B
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_chunked_pipeline.py tests/test_gpu_parity.py -m gpu -q -k "persistent or contract or fixed_iterations" --maxfail=8 > gpurun_out/r02_c2_pytest_z.log 2>&1; tail -15 gpurun_out/r02_c2_pytest_z.log
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r02_c2_pytest_mgpu.log 2>&1; tail -30 gpurun_out/r02_c2_pytest_mgpu.log
B="--no-cpu-baseline --no-extras --steps 10 --warmup 3"
for v in 2 1; do EVP_ZKERNEL=$v timeout 300 python bench.py --grid 256x256x512 $B > gpurun_out/r02_c2_1gpu_nz512_zk$v.json 2>gpurun_out/r02_c2_1gpu_nz512_zk$v.err; done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_c2_2gpu.json 2> gpurun_out/r02_c2_2gpu.err
EVP_ZKERNEL=1 timeout 600 $TR bench.py --gpus 2 $B > gpurun_out/r02_c2_2gpu_zk1.json 2> gpurun_out/r02_c2_2gpu_zk1.err
EVP_P2P_WAIT=read timeout 600 $TR bench.py --gpus 2 $B > gpurun_out/r02_c2_2gpu_waitread.json 2> gpurun_out/r02_c2_2gpu_waitread.err
EVP_TRANSPORT=nccl timeout 600 $TR bench.py --gpus 2 $B > gpurun_out/r02_c2_2gpu_nccl.json 2> gpurun_out/r02_c2_2gpu_nccl.err
timeout 600 $TR bench.py --gpus 2 --decomp pencil $B > gpurun_out/r02_c2_2gpu_pencil.json 2> gpurun_out/r02_c2_2gpu_pencil.err
for f in gpurun_out/r02_c2_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']), [(k['name'], k['ms']) for k in d.get('kernels',[])], 'exch', d.get('exchange_ms'), d.get('parity_check',{}).get('max_rel_diff_vs_cpu_oracle'))
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
