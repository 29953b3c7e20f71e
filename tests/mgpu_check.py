"""Run under torchrun (one rank per GPU): the N-rank decomposed solve (z-slabs or pencils, either transport) against the
CPU ORACLE on the full grid, same fixed iteration schedule (SURVEY.md §4 "1 GPU vs 2/4/8 GPU").

    torchrun ... tests/mgpu_check.py          MGPU_TRANSPORT=auto|nccl|p2p   MGPU_PY=1|2|4 (pencil rows)   MGPU_TMP=dir

Prints one line per deck and MGPU_OK / MGPU_FAIL.  `check_against_oracle()` is also what bench.py runs (small instance)
before timing a multi-rank configuration, so that every SCALE line carries a parity verdict.
The oracle is the checker here, never the thing measured."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lapx_b200 import api, build, distributed as dist, microstructure as ms  # noqa: E402

VOCE_HCP = [[5.0, 100.0, 5.0], [10.0, 200.0, 10.0], [20.0, 400.0, 20.0], [5.0, 50.0, 5.0]]
FIELDS = {"sig": api.FIELD_STRESS, "e": api.FIELD_STRAIN, "epsp": api.FIELD_PLASTIC_STRAIN, "crss": api.FIELD_CRSS}


def run(lib, host, grid, ng, hcp, dd, niter=6, nincs=2, seed=5):
    """Fixed schedule on `lib` (product library with `dd`, or the oracle with dd = None); fields of the local block."""
    ph = (ms.hcp_phase(host, with_twin=1, voce_mode=VOCE_HCP) if hcp else ms.fcc_phase(host, tau1=10.0, theta0=200.0, theta1=10.0))
    s = api.Solver(lib, grid, [ph], dist=dd)
    ids, grot = ms.voronoi_block(host, grid, ng, seed, s)
    s.set_microstructure(ids, None, ms.expand_rotations(ids, grot))
    s.set_reference_medium(None)
    s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=10**6, tol_newton=1e-9, newton_itmax=100)
    s.set_loading(api.Loading.uniaxial_tension(1.0))
    reps = []
    for inc in range(nincs):
        s.begin_increment(2e-4)
        for it in range(niter):
            r = s.equilibrium_iter()
            reps.append([r.err_stress, *r.savg, *r.emacro, r.newton_max, r.unconverged])
        s.end_increment()
    out = {k: s.get_field(f) for k, f in FIELDS.items()}
    out.update(c0=s.get_reference_medium(), reps=np.array(reps), block=np.array([s.y0, s.nyl, s.z0, s.nzl]),
               transport=s.transport() if dd is not None else "cpu", grain=s.get_field(api.FIELD_GRAIN))
    return out, s


def check_against_oracle(lib, td, world, rank, local, grid, ng, hcp, transport=api.TRANSPORT_AUTO, py=1, niter=6, nincs=2, tmp=None):
    """All ranks call this.  Returns (ok, info) on rank 0, (True, None) elsewhere."""
    import torch
    dd = dist.make_dist(lib, world, rank, local, td, transport=transport, py=py)
    part, s = run(lib, lib, grid, ng, hcp, dd, niter=niter, nincs=nincs)
    tmp = tmp or os.environ.get("MGPU_TMP", tempfile.gettempdir())
    np.savez(os.path.join(tmp, f"mgpu_{rank}.npz"), **part)
    torch.cuda.synchronize()   # this rank's own transposes (incl. the pulls already enqueued for a next iteration) have finished
    td.barrier()               # files written; nobody frees a buffer a peer may still be reading or writing
    s.close()
    ok, info = True, None
    if rank == 0:
        orc = api.load_library(build.build_oracle())
        orc.evp_oracle_set_threads(0)      # torchrun exports OMP_NUM_THREADS=1
        ref, so = run(orc, lib, grid, ng, hcp, None, niter=niter, nincs=nincs)
        so.close()
        nx, ny, nz = grid
        worst = {}
        parts = [np.load(os.path.join(tmp, f"mgpu_{r}.npz")) for r in range(world)]
        for k in list(FIELDS) + ["grain"]:
            full = np.zeros_like(ref[k])
            for p in parts:
                y0, nyl, z0, nzl = (int(v) for v in p["block"])
                full[:, z0:z0 + nzl, y0:y0 + nyl, :] = p[k]
            if k == "grain":
                worst[k] = 0.0 if np.array_equal(full, ref[k]) else 1.0       # indexing is bit exact
            else:
                worst[k] = float(np.abs(full - ref[k]).max() / np.abs(ref[k]).max())
        worst["reports"] = float(np.abs(part["reps"] - ref["reps"]).max() / np.abs(ref["reps"]).max())
        worst["c0"] = float(np.abs(part["c0"] - ref["c0"]).max() / np.abs(ref["c0"]).max())
        ok = all(v < 1e-8 for v in worst.values())
        info = {"grid": list(grid), "ranks": world, "py": py, "transport": str(part["transport"]), "hcp": bool(hcp),
                "iterations": niter * nincs, "max_rel_diff_vs_cpu_oracle": worst, "tolerance": 1e-8, "ok": bool(ok)}
    flag = torch.tensor([1 if ok else 0], device="cuda")
    td.broadcast(flag, 0)
    td.barrier()
    return bool(flag.item()), info


def main():
    import torch
    import torch.distributed as td
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = api.load_product()
    tr = {"auto": api.TRANSPORT_AUTO, "nccl": api.TRANSPORT_NCCL, "p2p": api.TRANSPORT_P2P}[os.environ.get("MGPU_TRANSPORT", "auto")]
    py = int(os.environ.get("MGPU_PY", "1"))
    ok = True
    # nz = 128 exercises the persistent radix-16 z kernel, nz = 512 the nz = 512 kernel, (64,16,32) the small one-shot kernels
    for grid, ng, hcp in [((32, 32, 64), 40, False), ((64, 16, 32), 25, True), ((16, 64, 128), 60, False), ((16, 16, 512), 40, False)]:
        if grid[1] % max(py, 1) or grid[1] % (world // py) or grid[2] % (world // py):
            continue
        good, info = check_against_oracle(lib, td, world, rank, local, grid, ng, hcp, transport=tr, py=py)
        if rank == 0:
            print("MGPU", info)
        ok &= good
    if rank == 0:
        print("MGPU_OK" if ok else "MGPU_FAIL")
    td.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
