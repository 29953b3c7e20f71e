"""CPU: known-answer tests that anchor the oracle to physics, since no reference
implementation is available to pin it (PARITY UNPINNED)."""
import numpy as np
import pytest
import scipy.fft as sfft

from common import make_polycrystal, rel_err
from lapx_b200 import api, microstructure as ms


def test_fft_matches_scipy(oracle_lib, product_lib):
    rng = np.random.default_rng(0)
    for grid in [(8, 8, 8), (12, 10, 6), (16, 8, 32), (9, 15, 5)]:
        s, ids, grot = make_polycrystal(oracle_lib, product_lib, grid, 4)
        nx, ny, nz = grid
        f = rng.normal(size=(6, nz, ny, nx))
        s.set_field(api.FIELD_STRESS, f)
        for comp in (0, 3, 5):
            spec = s.debug_spectrum(comp)
            ref = sfft.rfftn(f[comp], axes=(0, 1, 2))
            assert np.abs(spec - ref).max() < 1e-12 * np.abs(ref).max()


def test_single_crystal_uniform_and_analytic(oracle_lib, product_lib):
    """One grain: fields stay uniform, the iteration converges at once, and with the tensile axis along
    [001] of a cube-oriented FCC crystal the stress follows the scalar power law analytically."""
    ph = ms.fcc_phase(product_lib, gamma0=1.0, nrate=10.0, tau0=16.0)
    n = 8
    s = api.Solver(oracle_lib, (n, n, n), [ph])
    ids = np.zeros((n, n, n), np.int32)
    rot9 = ms.expand_rotations(ids, np.eye(3)[None])
    s.set_microstructure(ids, None, rot9)
    s.set_reference_medium(None)
    s.set_control(tol_stress=1e-10, tol_strain=1e-10, itmax=200, tol_newton=1e-12, newton_itmax=200)
    s.set_loading(api.Loading.uniaxial_tension(1.0))
    dt = 1e-4
    for inc in range(12):
        rep = s.step(dt)
        assert rep.converged
    sig = s.get_field(api.FIELD_STRESS)
    assert np.abs(sig - sig.reshape(6, -1)[:, :1].reshape(6, 1, 1, 1)).max() < 1e-9 * np.abs(sig).max()
    s33 = sig[2].mean()
    assert np.abs(sig[[0, 1, 3, 4, 5]]).max() < 1e-8 * s33
    # analytic: 8 active systems with Schmid factor 1/sqrt6; eps_p33_dot = 8 * m33 * g0 (m33 s/tau)^n, m33 = 1/sqrt6
    # total strain rate 1 = s33_dot / E001 + eps_p33_dot  -> implicit Euler gives the same s33 as the solver
    m33 = 1 / np.sqrt(6)
    c11, c12 = ms.CU_C11, ms.CU_C12
    E001 = (c11 - c12) * (c11 + 2 * c12) / (c11 + c12)
    sa, epa = 0.0, 0.0
    for inc in range(12):
        # solve  sa_new/E001 + epa + dt*8*m33*(m33*sa_new/16)^10 = (inc+1)*dt   (Newton)
        x = sa
        for _ in range(100):
            f = x / E001 + epa + dt * 8 * m33 * (m33 * x / 16.0) ** 10 - (inc + 1) * dt
            df = 1 / E001 + dt * 8 * m33 * 10 * (m33 * x / 16.0) ** 9 * m33 / 16.0
            x -= f / df
        sa = x
        epa += dt * 8 * m33 * (m33 * sa / 16.0) ** 10
    assert abs(s33 - sa) < 1e-8 * sa


def test_laminate_traction_and_compatibility(oracle_lib, product_lib):
    """Two-phase laminate normal to x (elastic-dominated step): sigma_1j continuous, in-plane e uniform."""
    ph = ms.fcc_phase(product_lib, tau0=1e6)      # effectively elastic
    nx, ny, nz = 16, 8, 8
    s = api.Solver(oracle_lib, (nx, ny, nz), [ph])
    ids = np.zeros((nz, ny, nx), np.int32)
    ids[:, :, nx // 2:] = 1
    rng = np.random.default_rng(3)
    grot = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(2)])
    grot *= np.sign(np.linalg.det(grot))[:, None, None]
    s.set_microstructure(ids, None, ms.expand_rotations(ids, grot))
    s.set_reference_medium(None)
    s.set_control(tol_stress=1e-11, tol_strain=1e-11, itmax=2000)
    D = np.array([[1.0, 0.2, 0.1], [0.2, -0.3, 0.0], [0.1, 0.0, -0.4]])
    s.set_loading(api.Loading.strain_rate(D))
    rep = s.step(1e-4)
    assert rep.converged
    sig = s.get_field(api.FIELD_STRESS)
    e = s.get_field(api.FIELD_STRAIN)
    # the laminate solution is piecewise uniform
    for f in (sig, e):
        for half in (slice(0, nx // 2), slice(nx // 2, nx)):
            blk = f[:, :, :, half]
            assert np.abs(blk - blk.mean(axis=(1, 2, 3), keepdims=True)).max() < 2e-7 * np.abs(f).max()
    left, right = sig[:, 0, 0, 1], sig[:, 0, 0, nx - 2]
    for c in (0, 4, 5):        # tractions on the x-plane: s11, s13, s12
        assert abs(left[c] - right[c]) < 1e-6 * np.abs(sig).max()
    el, er = e[:, 0, 0, 1], e[:, 0, 0, nx - 2]
    for c in (1, 2, 3):        # in-plane strains e22, e33, e23
        assert abs(el[c] - er[c]) < 1e-6 * np.abs(e).max()
    assert rel_err(e.reshape(6, -1).mean(axis=1), 1e-4 * np.array([1.0, -0.3, -0.4, 0.0, 0.1, 0.2])) < 1e-9


def test_green_operator_projector_identities(oracle_lib, product_lib):
    """Gamma applied to C0:grad_sym(u) of a periodic displacement returns that compatible strain, and
    applied to a divergence-free stress returns zero (both away from the Nyquist planes)."""
    n = 12  # not a power of two, and even -> Nyquist planes exist; use band-limited fields
    s, ids, grot = make_polycrystal(oracle_lib, product_lib, (n, n, n), 5)
    c0 = s.get_reference_medium()
    s.set_loading(api.Loading.strain_rate(np.zeros((3, 3))))
    x = np.arange(n) * 2 * np.pi / n
    Z, Y, X = np.meshgrid(x, x, x, indexing="ij")
    # displacement u = (sin(X+2Y), cos(2Z - X), sin(Y+Z)) -> strain (tensor components)
    e11 = np.cos(X + 2 * Y)
    e22 = np.zeros_like(X)
    e33 = np.cos(Y + Z)
    e23 = 0.5 * (-2 * np.sin(2 * Z - X) * 1.0 + np.cos(Y + Z))      # (du2/dz + du3/dy)/2
    e13 = np.zeros_like(X)
    e12 = 0.5 * (2 * np.cos(X + 2 * Y) + np.sin(2 * Z - X))          # (du1/dy + du2/dx)/2
    eps = np.stack([e11, e22, e33, e23, e13, e12])
    W = np.array([1, 1, 1, 2, 2, 2.0])
    sig = np.einsum("ab,b...->a...", c0 * W[None, :], eps)
    s.set_field(api.FIELD_STRESS, sig)
    s.set_field(api.FIELD_STRAIN, np.zeros_like(eps))
    s.begin_increment(1.0)
    s.op_green()                    # e <- 0 - Gamma*(C0:eps) = -eps
    got = -s.get_field(api.FIELD_STRAIN)
    assert np.abs(got - eps).max() < 1e-12 * np.abs(eps).max()


def _voce(ph, m, G):
    t0, t1, h0, h1 = ph.tau0[m], ph.tau1[m], ph.theta0[m], ph.theta1[m]
    return t0 + (t1 + h1 * G) * (1.0 - np.exp(-G * abs(h0 / t1)))


@pytest.mark.parametrize("hcp", [False, True])
def test_voce_hardening_follows_the_master_curve(hcp, oracle_lib, product_lib):
    """Extended Voce with all latent coefficients equal to 1: whatever systems are active, d tau_s = [tau^(G + dG) - tau^(G)] * sum_s'
    |d gamma_s'| / dG = tau^(G + dG) - tau^(G), so after every increment the CRSS of EVERY system of a voxel sits on the master
    curve of its mode at that voxel's accumulated shear: crss_s(x) = tau^_mode(s)(Gamma(x)) (Tome et al. 1984)."""
    voce = [[5.0, 100.0, 5.0], [10.0, 200.0, 10.0], [20.0, 400.0, 20.0], [5.0, 50.0, 5.0]]
    ph = (ms.hcp_phase(product_lib, with_twin=1, nrate=10.0, voce_mode=voce) if hcp else
          ms.fcc_phase(product_lib, nrate=10.0, tau0=16.0, tau1=10.0, theta0=200.0, theta1=10.0))
    s, ids, grot = make_polycrystal(oracle_lib, product_lib, (8, 8, 8), 6, seed=3, phase=ph)
    s.set_control(tol_stress=1e-6, tol_strain=1e-6, itmax=60, tol_newton=1e-10, newton_itmax=200)
    s.set_loading(api.Loading.plane_strain_compression(1.0))
    for inc in range(6):
        s.step(5e-4)
    crss, G = s.get_field(api.FIELD_CRSS), s.get_field(api.FIELD_GAMMA_ACC)[0]
    assert G.min() > 0 and G.max() > 1e-4                     # plastic everywhere by now
    for q in range(ph.nsys):
        ref = _voce(ph, ph.mode[q], G)
        assert np.abs(crss[q] - ref).max() < 1e-10 * ref.max(), q
    assert crss.min() > min(ph.tau0[m] for m in range(ph.nmodes))          # it did harden


def test_rate_sensitivity_saturation_stress(oracle_lib, product_lib):
    """[001] tension of a cube-oriented FCC crystal without hardening: the flow stress saturates where the plastic strain rate equals
    the imposed one, 8 m g0 (m s / tau)^n = rate with m = 1/sqrt 6, i.e. s_sat = (tau / m) (rate / (8 m g0))^(1/n); a ten times
    faster test raises it by 10^(1/n)."""
    n_exp, tau, g0, m = 10.0, 16.0, 1.0, 1 / np.sqrt(6)
    sat = {}
    for rate in (1.0, 10.0):
        ph = ms.fcc_phase(product_lib, gamma0=g0, nrate=n_exp, tau0=tau)
        s = api.Solver(oracle_lib, (8, 8, 8), [ph])
        ids = np.zeros((8, 8, 8), np.int32)
        s.set_microstructure(ids, None, ms.expand_rotations(ids, np.eye(3)[None]))
        s.set_reference_medium(None)
        s.set_control(tol_stress=1e-10, tol_strain=1e-10, itmax=200, tol_newton=1e-12, newton_itmax=200)
        s.set_loading(api.Loading.uniaxial_tension(rate))
        for inc in range(80):
            rep = s.step(1e-4 / rate)                       # same strain per increment at both rates
            assert rep.converged
        sat[rate] = rep.savg[2]
        analytic = (tau / m) * (rate / (8 * m * g0)) ** (1 / n_exp)
        assert abs(sat[rate] - analytic) < 1e-6 * analytic, (rate, sat[rate], analytic)
    assert abs(sat[10.0] / sat[1.0] - 10 ** (1 / n_exp)) < 1e-6


def test_elastic_crystal_rotates_with_the_material_spin(oracle_lib, product_lib):
    """Simple shear L12 = g of a single crystal that cannot slip (huge CRSS), texture update on: no plastic spin and no local (FFT)
    spin, so the lattice follows the applied spin W = (L - L^T)/2, R(t) = exp(W t) R0: a rotation by -g t / 2 about x3."""
    ph = ms.fcc_phase(product_lib, tau0=1e9)
    s = api.Solver(oracle_lib, (8, 8, 8), [ph])
    ids = np.zeros((8, 8, 8), np.int32)
    rng = np.random.default_rng(1)
    R0 = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    R0 *= np.sign(np.linalg.det(R0))
    s.set_microstructure(ids, None, ms.expand_rotations(ids, R0[None]))
    s.set_reference_medium(None)
    s.set_control(tol_stress=1e-10, tol_strain=1e-10, itmax=100, tol_newton=1e-12, newton_itmax=100, update_texture=1)
    g, dt, ninc = 0.8, 5e-3, 10
    L = np.zeros((3, 3))
    L[0, 1] = g
    s.set_loading(api.Loading.strain_rate(L))
    for inc in range(ninc):
        s.step(dt)
    th = -0.5 * g * dt * ninc
    Q = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    rot = s.get_field(api.FIELD_ROTATION).reshape(9, -1)
    assert np.abs(rot - (Q @ R0).reshape(9, 1)).max() < 1e-12
    assert np.abs(s.get_field(api.FIELD_LOCAL_ROTATION)).max() < 1e-14


def test_ptr_reorientation_is_the_85_degree_tensile_twin(oracle_lib, product_lib):
    """PTR reorientation of a {10-12} tensile twin in Zr (c/a = 1.594): the voxel takes the twin orientation R (2 n n^T - I), a rotation
    by 180 degrees about the twin-plane normal, which tilts the c axis by 2 atan((c/a)/sqrt 3) = 85.2 degrees (the textbook figure)."""
    import sys
    from common import GOLDEN
    sys.path.insert(0, GOLDEN)
    from common_golden import twin_phase
    ph = twin_phase(product_lib)
    grid = (8, 8, 8)
    ids, grot = ms.voronoi(product_lib, grid, 5, 11)
    s = api.Solver(oracle_lib, grid, [ph])
    s.set_microstructure(ids, None, ms.expand_rotations(ids, grot))
    s.set_reference_medium(None)
    s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=8, tol_newton=1e-9, newton_itmax=100, update_texture=0, update_twinning=1)
    s.set_loading(api.Loading.strain_rate(np.diag([-0.5, -0.5, 1.0])))
    r0 = s.get_field(api.FIELD_ROTATION).reshape(3, 3, -1)
    total = 0
    for inc in range(4):
        total += s.step(1e-3).reoriented
    tw = s.get_field(api.FIELD_TWINNED).reshape(-1).astype(bool)
    assert total == tw.sum() >= 1
    r1 = s.get_field(api.FIELD_ROTATION).reshape(3, 3, -1)
    assert np.array_equal(r1[:, :, ~tw], r0[:, :, ~tw])                      # texture update off: only twinned voxels changed
    c0, c1 = r0[:, 2, tw], r1[:, 2, tw]                                      # crystal c axis in the sample frame = third column of R
    ang = np.degrees(np.arccos(np.clip(np.einsum("iv,iv->v", c0, c1), -1, 1)))
    expect = np.degrees(2 * np.arctan(ms.ZR_COVERA / np.sqrt(3.0)))
    assert abs(expect - 85.2) < 0.1
    assert np.abs(ang - expect).max() < 1e-9
    for v in np.where(tw)[0]:                                                # proper rotation by pi: Q = R0^T R1 symmetric, trace -1
        Q = r0[:, :, v].T @ r1[:, :, v]
        assert np.abs(Q - Q.T).max() < 1e-12 and abs(np.trace(Q) + 1) < 1e-12 and abs(np.linalg.det(Q) - 1) < 1e-12
