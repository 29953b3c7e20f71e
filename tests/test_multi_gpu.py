"""GPU (-m gpu): N-rank decomposed runs equal the CPU oracle on the full grid (tests/mgpu_check.py under torchrun):
z-slabs with the peer-memory transport and with the NCCL transport, and pencils (py = 2).
Skipped when fewer than two GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("transport,py", [("p2p", 1), ("nccl", 1), ("nccl", 2)])
def test_ranks_match_oracle(transport, py, tmp_path):
    env = dict(os.environ, MGPU_TMP=str(tmp_path), MGPU_TRANSPORT=transport, MGPU_PY=str(py))
    n = min(torch.cuda.device_count(), 4)
    if n == 3:
        n = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, env=env, timeout=900)
    assert "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
