#!/usr/bin/env python
"""bench.py — voxel-updates/s per EVPFFT equilibrium iteration (fp64), BASELINE.json "metric".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one evp_equilibrium_iter over the whole grid.
Workload at N GPUs (weak scaling, 256^3 voxels per GPU, slab decomposition over z):
    N=1 256^3 | N=2 256x256x512 | N=4 256x512x512 | N=8 512^3   (FCC Cu Voronoi polycrystal, EVP tension)
The 256^3 / 512^3 grids are the ones the metric is quoted on; they are far larger than L2 (126 MB), so no
L2 flush is needed between timed iterations.

Printed keys beyond the base contract:
  roofline     dominant kernel, algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json hbm_gbs
  kernels      per-kernel device ms, algorithmic GB/s and fraction of the measured HBM peak
  cpu_baseline the CPU oracle (oracle/libevp_oracle.so, "port": the reference mount has no source) on the host cores
  e2e          same metric through the C-ABI with per-step host<->device traffic and host sync
The CPU oracle is only ever timed as a baseline / run as `--impl reference`; it is never on the product path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from lapx_b200 import api, microstructure as ms  # noqa: E402

METRIC = "voxel-updates/s per equilibrium iter (fp64)"
UNIT = "voxel-updates/s"
GRIDS = {1: (256, 256, 256), 2: (256, 256, 512), 4: (256, 512, 512), 8: (512, 512, 512)}
DT = 2e-4
PRE_ITERS = 10    # untimed iterations before the warm-up: the timed ones are mid-increment
TOL_NEWTON = float(os.environ.get("EVP_TOL_NEWTON", "1e-6"))   # library default; accepted iterate is accurate to ~tol^2


def phase_for(lib, workload):
    if workload == "hcp":
        return ms.hcp_phase(lib, with_twin=1, nrate=10.0,
                            voce_mode=[[5.0, 100.0, 5.0], [10.0, 200.0, 10.0], [20.0, 400.0, 20.0], [5.0, 50.0, 5.0]]), 24
    return ms.fcc_phase(lib, gamma0=1.0, nrate=10.0, tau0=16.0, tau1=10.0, theta0=200.0, theta1=10.0), 12


def algorithmic_bytes(grid, nsys):
    """Per-iteration algorithmic HBM bytes of each kernel."""
    nx, ny, nz = grid
    N = nx * ny * nz
    Nc = (nx // 2 + 1) * ny * nz
    return {
        "x_fwd": 48 * N + 96 * Nc,          # read 6 real fields, write 6 half spectra
        "y_fwd": 192 * Nc,
        "z_fused": 192 * Nc,
        "y_inv": 192 * Nc,
        "x_inv_update": 96 * Nc + 96 * N,   # read spectra, read+write e
        "constitutive": (24 + nsys) * 8 * N + 4 * N,   # sig read+write, e, eps_p, 1/tau_c, orientation class id
    }


KNAMES = ["x_fwd", "y_fwd", "z_fused", "y_inv", "x_inv_update", "constitutive"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        pmax = max(float(r[3]) for r in rows)
        busy = [r for r in rows if float(r[3]) >= 0.5 * pmax] or rows      # samples taken under load
        sm = sorted(float(r[1]) for r in busy)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(float(r[3]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for k, nme in enumerate(names):
                if r[5 + k].strip().lower().startswith("active"):
                    reasons.add(nme)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(rows)
        return out


def build_solver(lib, hostlib, grid, ngrains, workload, dist=None, z0=0, nzl=None):
    ph, nsys = phase_for(hostlib, workload)
    s = api.Solver(lib, grid, [ph], dist=dist)
    ids, grot = ms.voronoi(hostlib, grid, ngrains, 0, z0=s.z0, nzl=s.nzl)
    rot9 = ms.expand_rotations(ids, grot)
    t0 = time.perf_counter()
    s.set_microstructure(ids, None, rot9)
    t_up = time.perf_counter() - t0
    s._bench_micro = (ids, rot9)   # kept so that every leg of the bench can start from the same initial state
    up_bytes = ids.nbytes + rot9.nbytes
    s.set_reference_medium(None)
    s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=10**6, tol_newton=TOL_NEWTON, newton_itmax=100)
    ld = api.Loading.uniaxial_tension(1.0)
    s.set_loading(ld)
    return s, ld, nsys, t_up, up_bytes


def cpu_baseline(hostlib, workload, seconds_budget=20.0):
    """Oracle on the host cores, bounded sample: 64^3 (same material/BCs), mid-run iterations."""
    from lapx_b200 import build as b
    orc = api.load_library(b.build_oracle())
    grid = (64, 64, 64)
    s, ld, nsys, _, _ = build_solver(orc, hostlib, grid, 200, workload)
    s.begin_increment(DT)
    for _ in range(3):
        s.equilibrium_iter()
    t0 = time.perf_counter()
    n = 0
    while True:
        s.equilibrium_iter()
        n += 1
        el = time.perf_counter() - t0
        if el > seconds_budget or n >= 40:
            break
    cores = orc.evp_oracle_threads() if hasattr(orc, "evp_oracle_threads") else os.cpu_count()
    return {"value": grid[0] * grid[1] * grid[2] * n / el, "unit": UNIT, "cores": int(cores), "kind": "port",
            "sample": f"{n} mid-increment iterations of a 64^3 {workload.upper()} 200-grain polycrystal "
                      f"(same material, BCs and tolerances), OpenMP oracle incl. its own mixed-radix FFT"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lapx_b200 import build as b
    hostlib = api.load_product()
    orc = api.load_library(b.build_oracle())
    grid = (64, 64, 64)
    s, ld, nsys, _, _ = build_solver(orc, hostlib, grid, 200, args.workload)
    s.begin_increment(DT)
    for _ in range(args.warmup):
        s.equilibrium_iter()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.equilibrium_iter()
    el = time.perf_counter() - t0
    val = grid[0] * grid[1] * grid[2] * args.steps / el
    cores = int(orc.evp_oracle_threads())
    cb = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"each step = one iteration of a 64^3 {args.workload.upper()} 200-grain polycrystal (bounded sample of the "
                    f"{'x'.join(map(str, GRIDS[args.gpus]))} workload; the metric is size-normalised)"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload} Voronoi polycrystal, EVP uniaxial tension, CPU oracle (no reference source is mounted)",
                   "grid": list(grid)},
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="fcc", choices=["fcc", "hcp"])
    ap.add_argument("--grid", default=None, help="override, e.g. 128x128x128")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cufft", action="store_true", help="also time cuFFT (torch.fft) on the same 6 fields, as a comparison only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib = api.load_product()
    grid = GRIDS.get(args.gpus, GRIDS[1]) if args.grid is None else tuple(int(v) for v in args.grid.split("x"))
    dist = None
    if world > 1:
        import torch.distributed as td
        from lapx_b200 import distributed as evd
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = evd.make_dist(lib, world, rank, local, td)
    ngrains = 10000 if grid == (512, 512, 512) else max(50, int(round(grid[0] * grid[1] * grid[2] / 16777.216)))
    s, ld, nsys, t_up, up_bytes = build_solver(lib, lib, grid, ngrains, args.workload, dist=dist)
    N = grid[0] * grid[1] * grid[2]
    stream = torch.cuda.ExternalStream(s.stream())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            torch.cuda.synchronize()

    s.begin_increment(DT)
    clocks = ClockSampler(local)
    # get out of the cold start of the increment (Newton needs ~15 updates from sigma = 0 and the first
    # iterations are not representative of the bulk of an increment), then W warm-ups
    s.equilibrium_iters(PRE_ITERS)
    s.equilibrium_iters(args.warmup)

    # ---- device-timed value: K iterations back to back, inputs resident in HBM ----
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    rep = s.equilibrium_iters(args.steps)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_total = float(t.item())
    value = N * args.steps / (ms_total * 1e-3)

    # The e2e leg and the per-kernel leg repeat the value leg's iterations from the same initial state (microstructure
    # re-uploaded, fields reset, first increment, same pre-iterations and warm-up): late in an increment the warm-started
    # Newton solve needs one step instead of two, and a later increment a few voxels with three, so a leg that simply
    # followed the value leg would time a different workload.
    def restart_increment():
        s.set_microstructure(s._bench_micro[0], None, s._bench_micro[1])
        s.set_loading(ld)
        s.begin_increment(DT)
        s.equilibrium_iters(PRE_ITERS)
        s.equilibrium_iters(args.warmup)

    # ---- e2e through the C ABI: per step H2D of the boundary conditions, D2H of the report, host sync ----
    restart_increment()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.set_loading(ld)                      # H2D: 6x6 macro operator + imposed stress (host buffers)
        r = s.equilibrium_iter()               # D2H: iteration report (norms, <sigma>, E) after a stream sync
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_newton_mean = r.newton_mean
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        e2e_s = float(t.item())

    # ---- per-kernel device times (CUDA events recorded inside the library on its stream) ----
    restart_increment()
    s.set_profiling(1)
    kms = np.zeros(8)
    nprof = min(args.steps, 10)
    for _ in range(nprof):
        rk = s.equilibrium_iter()
        kms += s.last_kernel_ms()
    kms /= nprof
    prof_newton_mean = rk.newton_mean
    s.set_profiling(0)
    clk = clocks.stop()
    t0 = time.perf_counter()
    sig = s.get_field(api.FIELD_STRESS)
    t_down = time.perf_counter() - t0

    if world > 1:
        td.barrier()
    transport = "p2p" if lib.evp_transport(s.h) == 1 else "nccl"
    s.close()
    if world > 1:
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return
    hbm, peak_src = peaks()
    local_grid = (grid[0], grid[1], grid[2] // world)
    ab = algorithmic_bytes(local_grid, nsys)
    kern = []
    for i, k in enumerate(KNAMES):
        gbs = ab[k] / (kms[i] * 1e-3) / 1e9 if kms[i] > 0 else 0.0
        kern.append({"name": k, "ms": round(float(kms[i]), 4), "algorithmic_bytes": int(ab[k]), "gbs": round(gbs, 1),
                     "frac_hbm": round(gbs / hbm, 4)})
    dom = max(range(6), key=lambda i: kms[i])
    traffic, fp64 = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if list(tj.get("grid", [])) == list(local_grid) and KNAMES[dom] in tj and args.workload == "fcc":
            traffic = tj[KNAMES[dom]]["dram_bytes"]
            if "fp64_pipe_pct" in tj[KNAMES[dom]]:
                # second view of the same kernel (not timed here): share of cycles its fp64 pipe was busy in the committed
                # ncu capture, against the DFMA peak measured with scratch/dfma_bench.cu on this pool
                fp64 = {"pipe_active_pct_ncu": tj[KNAMES[dom]]["fp64_pipe_pct"], "peak_tflops_measured": tj["fp64_peak"]["tflops"],
                        "source": tj[KNAMES[dom]]["source"]}
    roof = {"kernel": KNAMES[dom], "bound": "hbm", "achieved": kern[dom]["gbs"], "peak": hbm, "unit": "GB/s",
            "frac": kern[dom]["frac_hbm"], "traffic": traffic, "peak_source": peak_src, "fp64": fp64,
            "note": "constitutive is fp64-pipe bound (DESIGN.md §4); its HBM fraction is reported for uniformity"}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{'x'.join(map(str, grid))} {args.workload.upper()} Voronoi polycrystal ({ngrains} grains), EVP "
                               f"uniaxial tension, mid-increment iterations, reference medium = Voigt average",
                   "grid": list(grid), "decomposition": "single GPU" if world == 1 else (f"z-slabs over {world} GPUs, FFT transposes = " + ("TMA stores into peer memory (CUDA IPC over NVLink) fused into the y/z passes" if transport == "p2p" else "NCCL all-to-all") + ", 4 pipelined z-chunks"),
                   "l2": "inputs larger than L2 (no flush needed)", "newton_mean": rep.newton_mean, "tol_newton": TOL_NEWTON,
                   "setup": {"h2d_bytes": int(up_bytes), "h2d_seconds": round(t_up, 4), "d2h_stress_bytes": int(sig.nbytes),
                             "d2h_seconds": round(t_down, 4)}},
        "roofline": roof, "kernels": kern, "kernels_newton_mean": prof_newton_mean,
        "exchange_ms": round(float(kms[6]), 4), "iter_ms_profiled": round(float(kms[7]), 4),
        "e2e": {"value": N * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 42 * 8, "d2h_bytes_per_step": 616,
                "newton_mean": e2e_newton_mean,   # same iterations of an increment as the value leg (config.newton_mean)
                "what": "evp_set_loading + evp_equilibrium_iter per step through the C ABI: BC upload, report download, host sync; "
                        "fields stay device resident by design (one-off transfer cost under config.setup)"},
        "gpu_launches": ((4 if world > 1 else 1) * 5 + 4) * args.steps,   # per iteration: 5 kernels per z-chunk + z pass + 2 reductions + macro
        "clocks": clk,
    }
    if args.cufft and world == 1:
        # comparison only (BASELINE.json north_star: "cuFFT timed only as a comparison"): 6 real fields, rfftn + irfftn
        x = torch.randn((6,) + tuple(reversed(grid)), dtype=torch.float64, device="cuda")
        for _ in range(2):
            y = torch.fft.irfftn(torch.fft.rfftn(x, dim=(1, 2, 3)), s=x.shape[1:], dim=(1, 2, 3))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = torch.fft.irfftn(torch.fft.rfftn(x, dim=(1, 2, 3)), s=x.shape[1:], dim=(1, 2, 3))
        e1.record()
        torch.cuda.synchronize()
        ours = sum(k["ms"] for k in kern if k["name"] != "constitutive")
        out["cufft_compare"] = {"cufft_rfftn_plus_irfftn_6_fields_ms": round(e0.elapsed_time(e1) / 5, 4),
                                "ours_fft_chain_ms": round(ours, 4),
                                "note": "ours includes the Green operator and the strain update; cuFFT figure is transforms only, out of place"}
        del x, y
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(lib, args.workload)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
