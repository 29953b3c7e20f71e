"""Run under torchrun: converge a few increments of the bench workload of N GPUs (512^3 FCC 10k grains on 8) to a
tolerance and print one JSON line with the stress-strain points, iteration counts and wall time per increment."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lapx_b200 import api, distributed as dist, microstructure as ms  # noqa: E402

GRIDS = {1: (256, 256, 256), 2: (256, 256, 512), 4: (256, 512, 512), 8: (512, 512, 512)}


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dd = None
    lib = api.load_product()
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
        dd = dist.make_dist(lib, world, rank, local, td)
    grid = GRIDS[world] if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split("x"))
    ng = 10000 if grid == (512, 512, 512) else max(50, grid[0] * grid[1] * grid[2] // 16777)
    tol = float(os.environ.get("EVP_TOL", "5e-5"))
    ph = ms.fcc_phase(lib, gamma0=1.0, nrate=10.0, tau0=16.0, tau1=10.0, theta0=200.0, theta1=10.0)
    s = api.Solver(lib, grid, [ph], dist=dd)
    ids, grot = ms.voronoi(lib, grid, ng, 0, z0=s.z0, nzl=s.nzl)
    s.set_microstructure(ids, None, ms.expand_rotations(ids, grot))
    s.set_reference_medium(None)
    s.set_control(tol_stress=tol, tol_strain=tol, itmax=300, itmin=2, tol_newton=1e-6, newton_itmax=100)
    s.set_loading(api.Loading.uniaxial_tension(1.0))
    incs = []
    for inc in range(int(os.environ.get("EVP_INCS", "3"))):
        t0 = time.perf_counter()
        r = s.step(2e-4)
        incs.append({"inc": inc + 1, "iters": r.iters, "converged": bool(r.converged), "err_stress": r.err_stress,
                     "err_strain": r.err_strain, "E33": r.emacro[2], "S33": r.savg[2], "S11": r.savg[0],
                     "epavg33": r.epavg[2], "seconds": round(time.perf_counter() - t0, 3)})
    if rank == 0:
        print(json.dumps({"grid": list(grid), "ranks": world, "grains": ng, "tol": tol, "increments": incs}))
    if world > 1:
        td.barrier()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
