"""CPU: the JSON contract of `bench.py --impl reference` (the arm the driver times beside the GPU arm), run on a tiny grid:
every key the driver reads is there, the CPU arm uses every online core whatever OMP_NUM_THREADS says, and ranks other than 0
print nothing under a multi-rank launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "16x16x16", "--steps", "2", "--warmup", "3", *args],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_line():
    out = _run({"OMP_NUM_THREADS": "1"})          # what torchrun exports: must not shrink the CPU arm
    d = json.loads(out.splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "voxel-updates/s" and d["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["config"]["grid"] == [16, 16, 16] and d["config"]["same_grid_as_gpu_arm"] is True
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == len(os.sched_getaffinity(0)) and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - 16**3) < 1e-6 * 16**3


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2") == ""
