"""Synthetic polycrystals for BASELINE.json's configs: periodic Voronoi grains with one random
orientation per grain.  Thin wrappers over the product library's host helpers
(lapx_b200/csrc/host_tables.cpp); nothing here needs a GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import Grid, Phase

# FCC Cu single-crystal constants in MPa (our choice; the contract gives none)
CU_C11, CU_C12, CU_C44 = 168400.0, 121400.0, 75400.0
# HCP Zr (room temperature) C11 C12 C13 C33 C44 in MPa, c/a
ZR_C5 = (143500.0, 72500.0, 65400.0, 164900.0, 32100.0)
ZR_COVERA = 1.594


def fcc_phase(lib, c11=CU_C11, c12=CU_C12, c44=CU_C44, gamma0=1.0, nrate=10.0, tau0=16.0,
              tau1=0.0, theta0=0.0, theta1=0.0) -> Phase:
    p = Phase()
    rc = lib.evp_phase_fcc(C.byref(p), c11, c12, c44, gamma0, nrate, tau0, tau1, theta0, theta1)
    assert rc == 0
    return p


def hcp_phase(lib, covera=ZR_COVERA, c5=ZR_C5, with_twin=1, gamma0=1.0, nrate=10.0,
              tau0_mode=(20.0, 100.0, 160.0, 80.0), voce_mode=None) -> Phase:
    p = Phase()
    c5a = np.ascontiguousarray(c5, np.float64)
    t0 = np.ascontiguousarray(tau0_mode, np.float64)
    vm = None if voce_mode is None else np.ascontiguousarray(voce_mode, np.float64).reshape(4, 3)
    rc = lib.evp_phase_hcp(C.byref(p), covera, c5a.ctypes.data_as(C.c_void_p), with_twin, gamma0, nrate,
                           t0.ctypes.data_as(C.c_void_p),
                           None if vm is None else vm.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return p


def voronoi(lib, grid, ngrains: int, seed: int = 0, z0: int = 0, nzl: int | None = None):
    """Grain ids of the slab [z0, z0+nzl) as int32 (nzl, ny, nx) and per-grain rotations (ngrains,3,3)."""
    nx, ny, nz = (int(v) for v in grid)
    nzl = nz - z0 if nzl is None else nzl
    g = Grid(nx, ny, nz, 1.0, 1.0, 1.0)
    ids = np.empty((nzl, ny, nx), np.int32)
    rot = np.empty((ngrains, 3, 3), np.float64)
    rc = lib.evp_voronoi(C.byref(g), ngrains, C.c_uint64(seed), z0, nzl, ids.ctypes.data_as(C.c_void_p),
                         rot.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return ids, rot


def voronoi_block(lib, grid, ngrains: int, seed: int, solver):
    """Grain ids of the local block of `solver` (slab or pencil) and the per-grain rotations."""
    ids, rot = voronoi(lib, grid, ngrains, seed, z0=solver.z0, nzl=solver.nzl)
    return np.ascontiguousarray(ids[:, solver.y0:solver.y0 + solver.nyl, :]), rot


def write_txt(lib, path: str, grain: np.ndarray, phase, rot9: np.ndarray):
    """Per-voxel text file "phi1 Phi phi2 i j k grain phase" (evp_write_microstructure_txt)."""
    nz, ny, nx = grain.shape
    g = Grid(nx, ny, nz, 1.0, 1.0, 1.0)
    gr = np.ascontiguousarray(grain, np.int32)
    ph = None if phase is None else np.ascontiguousarray(phase, np.int32)
    r = np.ascontiguousarray(rot9, np.float64)
    rc = lib.evp_write_microstructure_txt(str(path).encode(), C.byref(g), gr.ctypes.data_as(C.c_void_p),
                                          None if ph is None else ph.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise OSError(f"evp_write_microstructure_txt({path}) failed: {rc}")


def read_txt(lib, path: str, grid):
    """Read a per-voxel text file for `grid` = (nx, ny, nz): (grain[z,y,x], phase[z,y,x], rot9[9,z,y,x])."""
    nx, ny, nz = (int(v) for v in grid)
    g = Grid(nx, ny, nz, 1.0, 1.0, 1.0)
    grain = np.empty((nz, ny, nx), np.int32)
    phase = np.empty((nz, ny, nx), np.int32)
    rot9 = np.empty((9, nz, ny, nx), np.float64)
    rc = lib.evp_read_microstructure_txt(str(path).encode(), C.byref(g), grain.ctypes.data_as(C.c_void_p),
                                         phase.ctypes.data_as(C.c_void_p), rot9.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise OSError(f"evp_read_microstructure_txt({path}) failed: {rc} (missing file, index out of range, duplicate or missing voxels)")
    return grain, phase, rot9


def expand_rotations(grain: np.ndarray, grain_rot: np.ndarray) -> np.ndarray:
    """Per-voxel rotation field in the ABI layout [9][z][y][x] from per-grain matrices."""
    r = grain_rot.reshape(-1, 9)[grain.reshape(-1)]          # (nvox, 9)
    return np.ascontiguousarray(r.T).reshape((9,) + grain.shape)
