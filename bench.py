#!/usr/bin/env python
"""bench.py — voxel-updates/s per EVPFFT equilibrium iteration (fp64), BASELINE.json "metric".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload fcc|hcp] [--decomp slab|pencil]

One "step" = one evp_equilibrium_iter over the whole grid.
Workload at N GPUs (weak scaling, 256^3 voxels per GPU):
    N=1 256^3 | N=2 256x256x512 | N=4 256x512x512 | N=8 512^3   (FCC Cu Voronoi polycrystal, EVP uniaxial tension)
The 256^3 / 512^3 grids are the ones the metric is quoted on; they are far larger than L2 (126 MB), so no L2 flush is
needed between timed iterations.

Keys beyond the base contract:
  roofline      dominant kernel.  constitutive: fp64-bound -> algorithmic flops (static SASS count, profiles/k1_flops.json)
                / CUDA-event time vs the fp64 peak MEASURED IN THIS RUN (evp_debug_fp64_peak); the HBM view beside it.
  kernels       per-kernel device ms, algorithmic GB/s and fraction of the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle ("port": the reference mount has no source) on ALL host cores, SAME grid at N = 1
  e2e           same metric through the C ABI with per-step host<->device traffic and host sync
  increment     (N = 1) evp_step to tolerance + stress-field download through pinned staging: the per-increment end to end
  plastic       (N = 1) the same K iterations timed in the 6th increment (fully plastic), with newton_mean / newton_max
  parity_check  (N > 1) a small instance of the same decomposition/transport against the CPU oracle, run before timing
  config4_hcp   (N = 2, 4) BASELINE.json configs[3]: 256^3 HCP + twinning, timed on the same ranks
The CPU oracle is only ever timed as a baseline / run as `--impl reference` / used as the checker; never on the product path.
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
# library default.  tests/test_gpu_parity.py::test_bench_newton_tolerance_matches_tight_oracle holds the CUDA path at this
# tolerance to the oracle at 1e-12 within 1e-8 over the very schedule timed here.
TOL_NEWTON = float(os.environ.get("EVP_TOL_NEWTON", "1e-6"))
VOCE_HCP = [[5.0, 100.0, 5.0], [10.0, 200.0, 10.0], [20.0, 400.0, 20.0], [5.0, 50.0, 5.0]]
K1_KERNEL = {"fcc": "k_constitutive_p<12, 9, false, 4, 12, 1>", "hcp": "k_constitutive_p<24, 9, true, 3, 12, 2>"}


def phase_for(lib, workload):
    if workload == "hcp":
        return ms.hcp_phase(lib, with_twin=1, nrate=10.0, voce_mode=VOCE_HCP), 24
    return ms.fcc_phase(lib, gamma0=1.0, nrate=10.0, tau0=16.0, tau1=10.0, theta0=200.0, theta1=10.0), 12


def ngrains_for(grid):
    return 10000 if tuple(grid) == (512, 512, 512) else max(50, int(round(grid[0] * grid[1] * grid[2] / 16777.216)))


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
        "constitutive": (24 + nsys) * 8 * N + 4 * N,   # sig read+write, e, eps_p, rate factor, orientation class id
    }


KNAMES = ["x_fwd", "y_fwd", "z_fused", "y_inv", "x_inv_update", "constitutive"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def k1_flops(workload):
    """(F0, F1, source): algorithmic flops per voxel = F0 + F1 * newton_mean, static SASS count of the running kernel."""
    p = os.path.join(ROOT, "profiles", "k1_flops.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    k = d["kernels"].get(K1_KERNEL[workload])
    if not k:
        return None
    from lapx_b200 import build
    return k["F0"], k["F1"], {"file": "profiles/k1_flops.json", "kernel": K1_KERNEL[workload],
                              "counted_for_this_build": d.get("build_id") == "EVPSRC:" + build.source_id()}


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


def build_solver(lib, hostlib, grid, ngrains, workload, dist=None):
    ph, nsys = phase_for(hostlib, workload)
    s = api.Solver(lib, grid, [ph], dist=dist)
    ids, grot = ms.voronoi_block(hostlib, grid, ngrains, 0, s)
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


def load_oracle():
    """The CPU oracle with every online core (a launcher's OMP_NUM_THREADS=1 must not shrink the baseline)."""
    from lapx_b200 import build as b
    orc = api.load_library(b.build_oracle())
    cores = int(orc.evp_oracle_set_threads(0))
    return orc, cores


def scipy_fft_figure(grid):
    """The FFT share of one iteration with a library FFT on the same cores (the oracle's own FFT is a
    hand-written recursive one; this is the fairness figure): 6 components, rfftn + irfftn, all cores."""
    import scipy.fft as sfft
    nx, ny, nz = grid
    x = np.random.default_rng(0).normal(size=(6, nz, ny, nx))
    sfft.irfftn(sfft.rfftn(x[:1], axes=(1, 2, 3), workers=-1), s=(nz, ny, nx), axes=(1, 2, 3), workers=-1)
    t0 = time.perf_counter()
    y = sfft.irfftn(sfft.rfftn(x, axes=(1, 2, 3), workers=-1), s=(nz, ny, nx), axes=(1, 2, 3), workers=-1)
    el = time.perf_counter() - t0
    del x, y
    return round(1e3 * el, 1)


def cpu_baseline(hostlib, workload, grid, seconds_budget=20.0, max_iters=6):
    """Oracle on the host cores, SAME grid / material / BCs / tolerances as the GPU arm, mid-increment iterations."""
    orc, cores = load_oracle()
    s, ld, nsys, _, _ = build_solver(orc, hostlib, grid, ngrains_for(grid), workload)
    s.begin_increment(DT)
    s.equilibrium_iter()            # warm-up (first touch of the work arrays)
    t0 = time.perf_counter()
    n = 0
    while True:
        s.equilibrium_iter()
        n += 1
        el = time.perf_counter() - t0
        if el > seconds_budget or n >= max_iters:
            break
    s.close()
    return {"value": grid[0] * grid[1] * grid[2] * n / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} iterations (after 1 warm-up) of the SAME {'x'.join(map(str, grid))} {workload.upper()} workload "
                      f"(material, BCs, tolerances as the GPU arm); OpenMP oracle incl. its own mixed-radix FFT",
            "ms_per_step": round(1e3 * el / n, 1), "scipy_fft_6comp_rfftn_irfftn_ms": scipy_fft_figure(grid)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hostlib = api.load_product()
    orc, cores = load_oracle()
    # N = 1: the GPU arm's own grid.  N > 1: one rank's share of the weak-scaling workload (256^3 voxels), because the
    # CPU arm runs on rank 0's host alone and a 512^3 oracle iteration takes ~15 s.
    grid = GRIDS[1] if args.grid is None else tuple(int(v) for v in args.grid.split("x"))
    same = args.gpus == 1
    s, ld, nsys, _, _ = build_solver(orc, hostlib, grid, ngrains_for(grid), args.workload)
    s.begin_increment(DT)
    for _ in range(args.warmup):
        s.equilibrium_iter()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.equilibrium_iter()
    el = time.perf_counter() - t0
    val = grid[0] * grid[1] * grid[2] * args.steps / el
    what = ("the GPU arm's own workload" if same else
            f"one rank's share (256^3 voxels) of the {'x'.join(map(str, GRIDS.get(args.gpus, GRIDS[1])))} weak-scaling workload; the metric is size-normalised")
    cb = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"each step = one iteration of a {'x'.join(map(str, grid))} {args.workload.upper()} Voronoi polycrystal: {what}"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{'x'.join(map(str, grid))} {args.workload.upper()} Voronoi polycrystal ({ngrains_for(grid)} grains), EVP uniaxial "
                               f"tension, mid-increment iterations, reference medium = Voigt average; CPU oracle (no reference source is mounted)",
                   "grid": list(grid), "tol_newton": TOL_NEWTON, "same_grid_as_gpu_arm": same},
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
    ap.add_argument("--decomp", default="slab", choices=["slab", "pencil"], help="N > 1: z-slabs (default) or py x pz pencils")
    ap.add_argument("--py", type=int, default=0, help="pencil rows (default 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the increment / plastic / config-4 / parity legs")
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
    from lapx_b200 import build
    lib = api.load_product()
    build_id = lib.evp_build_id().decode()
    grid = GRIDS.get(args.gpus, GRIDS[1]) if args.grid is None else tuple(int(v) for v in args.grid.split("x"))
    dist = None
    py = 1
    td = None
    parity = None
    if world > 1:
        import torch.distributed as td
        from lapx_b200 import distributed as evd
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
        if args.decomp == "pencil":
            py = args.py if args.py > 1 else 2
        transport = api.TRANSPORT_NCCL if py > 1 else api.TRANSPORT_AUTO
        if not args.no_extras:
            # driver-visible multi-rank parity: a small instance of this decomposition / transport against the CPU oracle
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import mgpu_check
            ok, parity = mgpu_check.check_against_oracle(lib, td, world, rank, local, (64, 64, 128), 100, False, transport=transport, py=py,
                                                         niter=4, nincs=1)
            if not ok:
                if rank == 0:
                    print(json.dumps({"error": "multi-rank parity check against the CPU oracle FAILED; not timing", "parity_check": parity}))
                td.destroy_process_group()
                raise SystemExit(3)
        dist = evd.make_dist(lib, world, rank, local, td, transport=transport, py=py)
    ngrains = ngrains_for(grid)
    s, ld, nsys, t_up, up_bytes = build_solver(lib, lib, grid, ngrains, args.workload, dist=dist)
    N = grid[0] * grid[1] * grid[2]
    stream = torch.cuda.ExternalStream(s.stream())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            torch.cuda.synchronize()

    def maxrank(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    def timed_iters(k):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = s.launch_count()
        ev0.record(stream)
        rep = s.equilibrium_iters(k)
        ev1.record(stream)
        barrier()
        return maxrank(ev0.elapsed_time(ev1)), rep, s.launch_count() - l0

    s.begin_increment(DT)
    clocks = ClockSampler(local)
    # get out of the cold start of the increment (Newton needs ~15 updates from sigma = 0 and the first
    # iterations are not representative of the bulk of an increment), then W warm-ups
    s.equilibrium_iters(PRE_ITERS)
    s.equilibrium_iters(args.warmup)

    # ---- device-timed value: K iterations back to back, inputs resident in HBM ----
    ms_total, rep, launches = timed_iters(args.steps)
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
    e2e_s = maxrank(time.perf_counter() - t0)
    e2e_newton_mean = r.newton_mean

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
    fp64_peak = s.fp64_peak_tflops(5)          # measured now, on this GPU, on the library's stream

    extras = {}
    if not args.no_extras and world == 1:
        # ---- the per-increment end to end: evp_step to tolerance + download of the stress field (pinned staging) ----
        s.set_microstructure(s._bench_micro[0], None, s._bench_micro[1])
        s.set_loading(ld)
        s.set_control(tol_stress=5e-5, tol_strain=5e-5, itmax=200, tol_newton=TOL_NEWTON, newton_itmax=100)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.set_loading(ld)
        sr = s.step(DT)
        t_step = time.perf_counter() - t0
        sig = s.get_field(api.FIELD_STRESS)
        t_all = time.perf_counter() - t0
        extras["increment"] = {"what": "evp_step (begin + iterate to err <= 5e-5 + commit) and evp_get_field(STRESS) into a pageable host array",
                               "iters": sr.iters, "converged": sr.converged, "seconds_step": round(t_step, 4), "seconds_total": round(t_all, 4),
                               "d2h_bytes": int(sig.nbytes), "d2h_gbs": round(sig.nbytes / max(t_all - t_step, 1e-9) / 1e9, 2),
                               "value": N * sr.iters / t_all, "unit": UNIT}
        del sig
        # ---- a fully plastic increment: 4 more increments to tolerance, then the same timing protocol in the 6th ----
        for _ in range(4):
            sr = s.step(DT)
        s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=10**6, tol_newton=TOL_NEWTON, newton_itmax=100)
        s.begin_increment(DT)
        s.equilibrium_iters(PRE_ITERS)
        s.equilibrium_iters(args.warmup)
        ms_p, rp, _ = timed_iters(args.steps)
        e_mac, s_avg = s.get_macro()
        extras["plastic"] = {"what": f"{args.steps} mid-increment iterations of increment 6 (E33 = {e_mac[2]:.2e}, <s33> = {s_avg[2]:.1f} MPa)",
                             "ms_per_step": ms_p / args.steps, "value": N * args.steps / (ms_p * 1e-3), "unit": UNIT,
                             "newton_mean": rp.newton_mean, "newton_max": rp.newton_max, "unconverged": rp.unconverged}
    clk = clocks.stop()
    t0 = time.perf_counter()
    sig = s.get_field(api.FIELD_STRESS)
    t_down = time.perf_counter() - t0

    if world > 1:
        torch.cuda.synchronize()     # peers pull from / push into this rank's spectral buffers until its streams are idle
        td.barrier()
    transport = s.transport()
    s.close()
    del s

    # ---- BASELINE.json configs[3]: 256^3 HCP with twinning on 2-4 GPUs, timed on the same ranks ----
    if not args.no_extras and world in (2, 4) and args.workload == "fcc" and args.grid is None:
        td.barrier()
        g4 = (256, 256, 256)
        s4, ld4, nsys4, _, _ = build_solver(lib, lib, g4, 1000, "hcp", dist=evd.make_dist(lib, world, rank, local, td, transport=transport_id(transport), py=py))
        s4.begin_increment(DT)
        s4.equilibrium_iters(PRE_ITERS)
        s4.equilibrium_iters(args.warmup)
        torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
        st4 = torch.cuda.ExternalStream(s4.stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st4)
        r4 = s4.equilibrium_iters(args.steps)
        e1.record(st4)
        torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
        ms4 = maxrank(e0.elapsed_time(e1))
        extras["config4_hcp"] = {"workload": "256x256x256 HCP Zr (3 prismatic + 3 basal + 12 pyramidal<c+a> + 6 tensile twins), 1000 grains, "
                                             "EVP uniaxial tension, mid-increment iterations", "grid": list(g4), "n_gpus": world, "scaling": "strong",
                                 "ms_per_step": ms4 / args.steps, "value": 256**3 * args.steps / (ms4 * 1e-3), "unit": UNIT,
                                 "newton_mean": r4.newton_mean, "transport": s4.transport()}
        torch.cuda.synchronize()
        td.barrier()
        s4.close()
    if world > 1:
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return
    hbm, peak_src = peaks()
    local_grid = (grid[0], grid[1] // py, grid[2] // (world // py))
    ab = algorithmic_bytes(local_grid, nsys)
    kern = []
    for i, k in enumerate(KNAMES):
        gbs = ab[k] / (kms[i] * 1e-3) / 1e9 if kms[i] > 0 else 0.0
        kern.append({"name": k, "ms": round(float(kms[i]), 4), "algorithmic_bytes": int(ab[k]), "gbs": round(gbs, 1),
                     "frac_hbm": round(gbs / hbm, 4)})
    dom = max(range(6), key=lambda i: kms[i])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if list(tj.get("grid", [])) == list(local_grid) and KNAMES[dom] in tj and args.workload == "fcc":
            traffic = tj[KNAMES[dom]]["dram_bytes"]     # dram__bytes_read + write of one launch, ncu --set full capture (profiles/)
    nloc = local_grid[0] * local_grid[1] * local_grid[2]
    fl = k1_flops(args.workload)
    if KNAMES[dom] == "constitutive" and fl is not None:
        f0, f1, src = fl
        flops = (f0 + f1 * prof_newton_mean) * nloc
        tf = flops / (kms[5] * 1e-3) / 1e12
        roof = {"kernel": "constitutive", "bound": "fp64", "achieved": round(tf, 3), "peak": round(fp64_peak, 3), "unit": "TFLOP/s",
                "frac": round(tf / fp64_peak, 4), "traffic": traffic,
                "peak_source": "evp_debug_fp64_peak: dependent-chain DFMA microbenchmark, 32 warps/SM x 8 chains, measured in this run",
                "algorithmic_flops_per_voxel": round(f0 + f1 * prof_newton_mean, 1), "flops_source": src,
                "hbm_view": {"achieved_gbs": kern[5]["gbs"], "peak_gbs": hbm, "frac": kern[5]["frac_hbm"], "peak_source": peak_src,
                             "algorithmic_bytes": kern[5]["algorithmic_bytes"]}}
    else:
        roof = {"kernel": KNAMES[dom], "bound": "hbm", "achieved": kern[dom]["gbs"], "peak": hbm, "unit": "GB/s",
                "frac": kern[dom]["frac_hbm"], "traffic": traffic, "peak_source": peak_src,
                "note": ("multi-rank: this kernel's TMA stores / loads on peer memory are the FFT transpose over NVLink; its time includes the exchange"
                         if world > 1 and transport == "p2p" and KNAMES[dom] in ("y_inv", "y_fwd") else None)}
    decomp = ("single GPU" if world == 1 else
              (f"{py} x {world // py} pencils, 4 NCCL all-to-alls per iteration on a communication stream" if py > 1 else
               f"z-slabs over {world} GPUs, FFT transposes = " + ("peer memory over NVLink (CUDA IPC): forward = TMA stores into the peers' receive buffers "
                                                                  "fused into the forward y pass, way back = TMA loads out of the peers' z-pass buffers fused into "
                                                                  "the inverse y pass" if transport == "p2p" else "NCCL all-to-all") + ", pipelined over z-chunks"))
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{'x'.join(map(str, grid))} {args.workload.upper()} Voronoi polycrystal ({ngrains} grains), EVP "
                               f"uniaxial tension, mid-increment iterations, reference medium = Voigt average",
                   "grid": list(grid), "decomposition": decomp,
                   "l2": "inputs larger than L2 (no flush needed)", "newton_mean": rep.newton_mean, "tol_newton": TOL_NEWTON,
                   "setup": {"h2d_bytes": int(up_bytes), "h2d_seconds": round(t_up, 4), "h2d_gbs": round(up_bytes / t_up / 1e9, 2),
                             "d2h_stress_bytes": int(sig.nbytes), "d2h_seconds": round(t_down, 4), "d2h_gbs": round(sig.nbytes / t_down / 1e9, 2)}},
        "roofline": roof, "kernels": kern, "kernels_newton_mean": prof_newton_mean, "fp64_peak_tflops_measured": round(fp64_peak, 3),
        "exchange_ms": round(float(kms[6]), 4), "iter_ms_profiled": round(float(kms[7]), 4),
        "e2e": {"value": N * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 42 * 8, "d2h_bytes_per_step": 624,
                "newton_mean": e2e_newton_mean,   # same iterations of an increment as the value leg (config.newton_mean)
                "what": "evp_set_loading + evp_equilibrium_iter per step through the C ABI: BC upload, report download, host sync; "
                        "fields stay device resident by design (one-off transfer cost under config.setup; per-increment e2e under `increment`)"},
        "gpu_launches": int(launches),    # counted by the library (evp_launch_count) across the timed region
        "clocks": clk,
        "build_id": build_id, "build_id_matches_sources": build_id == "EVPSRC:" + build.source_id(),
    }
    out.update(extras)
    if parity is not None:
        out["parity_check"] = parity
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
        out["cpu_baseline"] = cpu_baseline(lib, args.workload, grid)
    print(json.dumps(out))


def transport_id(name):
    return api.TRANSPORT_P2P if name == "p2p" else api.TRANSPORT_NCCL


if __name__ == "__main__":
    main()
