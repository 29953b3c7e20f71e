"""GPU (-m gpu): 2-rank slab-decomposed run (NCCL all-to-all FFT transposes) equals the single-GPU run.
Skipped when fewer than two GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_ranks_match_one_rank(tmp_path):
    env = dict(os.environ, MGPU_TMP=str(tmp_path))
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
