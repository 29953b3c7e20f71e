"""Shared helpers for the parity tests: build the same synthetic case for any ABI implementation."""
import os

import numpy as np

from lapx_b200 import api, microstructure as ms

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_phase(host, hcp: bool, kind: int = 0):
    if kind in (2, 3):
        import sys
        sys.path.insert(0, GOLDEN)
        from common_golden import twin_phase, twin_phase_ratio
        return twin_phase(host) if kind == 2 else twin_phase_ratio(host)
    if hcp:
        return ms.hcp_phase(host, with_twin=1, nrate=10.0,
                            voce_mode=[[5.0, 100.0, 5.0], [10.0, 200.0, 10.0], [20.0, 400.0, 20.0], [5.0, 50.0, 5.0]])
    return ms.fcc_phase(host, gamma0=1.0, nrate=10.0, tau0=16.0, tau1=10.0, theta0=200.0, theta1=10.0)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def solver_from_golden(lib, host, g, c0="golden"):
    """Create a Solver on `lib` with the inputs stored in golden file `g` (phase tables from `host`)."""
    grid = tuple(int(v) for v in g["grid"])
    ph = golden_phase(host, bool(int(g["hcp"])), int(g["phase_kind"]) if "phase_kind" in g.files else 0)
    s = api.Solver(lib, grid, [ph])
    rot9 = ms.expand_rotations(g["grain"], g["grain_rot"])
    s.set_microstructure(g["grain"], None, rot9)
    s.set_reference_medium(g["c0_voigt"] if c0 == "golden" else None)
    s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=10**6, itmin=1, tol_newton=1e-9, newton_itmax=100,
                  update_texture=int(g["texture"]) if "texture" in g.files else 0,
                  update_twinning=int(g["twinning"]) if "twinning" in g.files else 0)
    s.set_loading(api.Loading(g["iudot"], g["udot"], g["iscau"], g["scau"]))
    return s


def run_golden_schedule(s, g, hook=None, step_reports=None):
    """Run the fixed iteration schedule of a golden file; returns the report table (same columns).
    `step_reports`: optional list that receives the evp_step_report of every end_increment."""
    rows = []
    dt = float(g["dt"])
    for inc in range(int(g["nincs"])):
        s.begin_increment(dt)
        for it in range(int(g["iters_per_inc"])):
            s.op_green()
            if hook:
                hook(s, inc, it, "green")
            r = s.op_constitutive()
            if hook:
                hook(s, inc, it, "const")
            rows.append([inc, it + 1, r.err_stress, r.err_strain, *r.savg, *r.emacro, r.newton_max, r.newton_mean])
        sr = s.end_increment()
        if step_reports is not None:
            step_reports.append(sr)
        if hook:
            hook(s, inc, -1, "end")
    return np.array(rows)


def rel_err(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    d = np.abs(a - b).max()
    return d / max(np.abs(b).max(), 1e-300)


def make_polycrystal(lib, host, grid, ngrains, seed=0, hcp=False, phase=None, c0=None):
    ph = phase if phase is not None else golden_phase(host, hcp)
    ids, grot = ms.voronoi(host, grid, ngrains, seed)
    s = api.Solver(lib, grid, [ph])
    s.set_microstructure(ids, None, ms.expand_rotations(ids, grot))
    s.set_reference_medium(c0)
    return s, ids, grot
