"""GPU (-m gpu): the z-chunked, pipelined iteration (used with ranks > 1 to overlap the FFT transposes) run on one
GPU with EVP_CHUNKS = 2/4 must reproduce the unchunked solve."""
import numpy as np
import pytest

from common import make_polycrystal, rel_err
from lapx_b200 import api

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("grid,hcp", [((8, 8, 8), False), ((32, 32, 32), False), ((16, 16, 64), True), ((128, 16, 32), False)])
def test_chunked_equals_unchunked(grid, hcp, product_lib, monkeypatch):
    outs = {}
    for chunks in (1, 2, 4):
        monkeypatch.setenv("EVP_CHUNKS", str(chunks))
        s, ids, grot = make_polycrystal(product_lib, product_lib, grid, 12, seed=3, hcp=hcp)
        s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=10**6, tol_newton=1e-9, newton_itmax=100)
        s.set_loading(api.Loading.uniaxial_tension(1.0))
        reps = []
        for inc in range(2):
            s.begin_increment(2e-4)
            for it in range(5):
                r = s.equilibrium_iter()
                reps.append([r.err_stress, r.err_strain, *r.savg, *r.emacro, r.newton_max, r.newton_mean])
            s.end_increment()
        outs[chunks] = (np.array(reps), s.get_field(api.FIELD_STRESS), s.get_field(api.FIELD_STRAIN), s.get_field(api.FIELD_CRSS))
    for chunks in (2, 4):
        for a, b in zip(outs[chunks], outs[1]):
            assert rel_err(a, b) < 1e-12, chunks


@pytest.mark.parametrize("grid", [(16, 16, 128), (32, 8, 256), (8, 64, 128), (256, 64, 256), (16, 16, 512), (64, 8, 512), (256, 64, 512)])
def test_persistent_z_kernel_equals_one_shot(grid, product_lib, monkeypatch):
    """k_zfused2 (nz = 128: persistent, radix-16), k_zfused4 (nz = 256) and k_zfused3 (nz = 512) (persistent, in-place passes,
    component slots pipelined by producer warps; the large grids give every block several tiles) vs k_zfused (one tile per
    block, radix-8 Stockham passes)."""
    outs = []
    variants = [(0, "0"), (4, "0")] + ([(0, "1")] if grid[2] == 256 else [])      # nz = 256: also the opt-in k_zfused4
    for flags, z4 in variants:
        monkeypatch.setenv("EVP_Z4", z4)
        s, ids, grot = make_polycrystal(product_lib, product_lib, grid, 12, seed=5)
        s.set_profiling(flags)
        s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=10**6, tol_newton=1e-9, newton_itmax=100)
        s.set_loading(api.Loading.uniaxial_tension(1.0))
        s.begin_increment(2e-4)
        for it in range(5):
            r = s.equilibrium_iter()
        outs.append((s.get_field(api.FIELD_STRESS), s.get_field(api.FIELD_STRAIN), np.array(r.savg[:])))
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert rel_err(a, b) < 1e-11


@pytest.mark.parametrize("hcp", [False, True, 2])
def test_constitutive_kernel_variants_agree(hcp, product_lib, monkeypatch):
    """k_constitutive_p (bulk-staged fast path; FCC: compile-time Schmid table) vs the same kernel with run-time tables
    (EVP_K1_FCC=0) vs the thread-loads kernel k_constitutive_t (EVP_K1_LEGACY=1) vs the fast path without bulk staging
    (EVP_K1_BULK=0: the per-thread loads partial blocks take): same Newton iteration, different arithmetic organisation
    (DESIGN.md section 4)."""
    outs = []
    for env in ({}, {"EVP_K1_FCC": "0"}, {"EVP_K1_LEGACY": "1"}, {"EVP_K1_BULK": "0"}):
        for k in ("EVP_K1_FCC", "EVP_K1_LEGACY", "EVP_K1_BULK"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ph = None
        if hcp == 2:   # 30 systems: HCP with tensile and compressive twins (fast path vs the generic thread-loads kernel)
            from lapx_b200 import microstructure as ms
            ph = ms.hcp_phase(product_lib, with_twin=2, nrate=10.0, voce_mode=[[5.0, 100.0, 5.0], [10.0, 200.0, 10.0], [20.0, 400.0, 20.0], [5.0, 50.0, 5.0]])
        s, ids, grot = make_polycrystal(product_lib, product_lib, (32, 16, 16), 12, seed=9, hcp=bool(hcp), phase=ph)
        s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=10**6, tol_newton=1e-9, newton_itmax=100)
        s.set_loading(api.Loading.uniaxial_tension(1.0))
        reps = []
        for inc in range(2):
            s.begin_increment(2e-4)
            for it in range(6):
                r = s.equilibrium_iter()
                reps.append([r.err_stress, r.err_strain, *r.savg, *r.emacro, r.newton_max])
            s.end_increment()
        outs.append((np.array(reps), s.get_field(api.FIELD_STRESS), s.get_field(api.FIELD_STRAIN), s.get_field(api.FIELD_CRSS)))
        s.close()
    for other in outs[1:]:
        for a, b in zip(other, outs[0]):
            assert rel_err(a, b) < 1e-10
