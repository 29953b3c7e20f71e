"""Run under torchrun (one rank per GPU): N-rank slab-decomposed solve vs the single-GPU solve of the
same deck, same iteration count (SURVEY.md §4 "1 GPU vs 2/4/8 GPU").  Prints MGPU_OK on success."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lapx_b200 import api, distributed as dist, microstructure as ms  # noqa: E402


def run(lib, grid, ng, hcp, dd, niter=6, nincs=2):
    ph = (ms.hcp_phase(lib, with_twin=1, voce_mode=[[5.0, 100.0, 5.0], [10.0, 200.0, 10.0], [20.0, 400.0, 20.0], [5.0, 50.0, 5.0]])
          if hcp else ms.fcc_phase(lib, tau1=10.0, theta0=200.0, theta1=10.0))
    s = api.Solver(lib, grid, [ph], dist=dd)
    ids, grot = ms.voronoi(lib, grid, ng, 5, z0=s.z0, nzl=s.nzl)
    s.set_microstructure(ids, None, ms.expand_rotations(ids, grot))
    s.set_reference_medium(None)
    s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=10**6, tol_newton=1e-9, newton_itmax=100)
    s.set_loading(api.Loading.uniaxial_tension(1.0))
    reps = []
    for inc in range(nincs):
        s.begin_increment(2e-4)
        for it in range(niter):
            r = s.equilibrium_iter()
            reps.append([r.err_stress, r.err_strain, *r.savg, *r.emacro, r.newton_max])
        s.end_increment()
    out = {"sig": s.get_field(api.FIELD_STRESS), "e": s.get_field(api.FIELD_STRAIN), "crss": s.get_field(api.FIELD_CRSS),
           "c0": s.get_reference_medium(), "reps": np.array(reps), "z0": s.z0, "nzl": s.nzl}
    if dd is not None:
        td.barrier()          # p2p transport: nobody frees a buffer a peer may still be writing to
    s.close()
    return out


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = api.load_product()
    tmp = os.environ.get("MGPU_TMP", tempfile.gettempdir())
    ok = True
    for grid, ng, hcp in [((32, 32, 64), 40, False), ((64, 16, 32), 25, True), ((16, 64, 128), 60, False)]:
        dd = dist.make_dist(lib, world, rank, local, td)
        part = run(lib, grid, ng, hcp, dd)
        np.savez(os.path.join(tmp, f"mgpu_{rank}.npz"), **part)
        td.barrier()
        if rank == 0:
            ref = run(lib, grid, ng, hcp, None)
            for k in ("sig", "e", "crss"):
                got = np.concatenate([np.load(os.path.join(tmp, f"mgpu_{r}.npz"))[k] for r in range(world)], axis=1)
                err = np.abs(got - ref[k]).max() / np.abs(ref[k]).max()
                print(f"grid {grid} hcp={hcp} {k}: max rel diff {world} ranks vs 1 rank = {err:.3e}")
                ok &= bool(err < 1e-10)
            rerr = np.abs(part["reps"] - ref["reps"]).max() / np.abs(ref["reps"]).max()
            cerr = np.abs(part["c0"] - ref["c0"]).max() / np.abs(ref["c0"]).max()
            print(f"  reports diff {rerr:.3e}, C0 diff {cerr:.3e}")
            ok &= bool(rerr < 1e-10 and cerr < 1e-12)
        td.barrier()
    if rank == 0:
        print("MGPU_OK" if ok else "MGPU_FAIL")
    td.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
