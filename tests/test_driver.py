"""The C++ host driver: deck parsing on CPU; on the GPU its stress-strain curve equals the one
obtained through the Python mirror of the C ABI for the same deck."""
import os
import subprocess

import numpy as np
import pytest

from lapx_b200 import api, build, microstructure as ms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "examples", "fcc32_tension.deck")


@pytest.fixture(scope="module")
def driver(product_lib):
    return build.build_driver()


def test_driver_parses_deck(driver):
    r = subprocess.run([driver, "--check", DECK], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "grid 32 32 32" in r.stdout and "voronoi 50 grains seed 0" in r.stdout
    assert "iudot 0 1 1 1 0 1 1 1 1" in r.stdout and "iscau 1 1 0 0 0 0" in r.stdout
    bad = subprocess.run([driver, "--check", os.path.join(ROOT, "README.md")], capture_output=True, text=True)
    assert bad.returncode != 0 and "unknown key" in bad.stderr


@pytest.mark.gpu
def test_driver_curve_matches_api(driver, product_lib, tmp_path):
    deck = tmp_path / "deck.txt"
    deck.write_text(open(DECK).read().replace("increments 5", "increments 3").replace("fcc32_tension_curve.txt", str(tmp_path / "curve.txt"))
                    + f"output_field stress {tmp_path / 'stress.bin'}\n")
    r = subprocess.run([driver, str(deck)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    curve = np.loadtxt(tmp_path / "curve.txt")
    assert curve.shape == (3, 24)
    # same deck through the Python mirror
    grid = (32, 32, 32)
    ph = ms.fcc_phase(product_lib, gamma0=1.0, nrate=10.0, tau0=16.0, tau1=10.0, theta0=200.0, theta1=10.0)
    ids, grot = ms.voronoi(product_lib, grid, 50, 0)
    s = api.Solver(product_lib, grid, [ph])
    s.set_microstructure(ids, None, ms.expand_rotations(ids, grot))
    s.set_reference_medium(None)
    s.set_control(tol_stress=1e-5, tol_strain=1e-5, itmax=200, itmin=1, tol_newton=1e-6, newton_itmax=100)
    s.set_loading(api.Loading.uniaxial_tension(1.0))
    for inc in range(3):
        rep = s.step(2e-4)
        assert int(curve[inc, 1]) == rep.iters and int(curve[inc, 2]) == rep.converged == 1
        assert np.allclose(curve[inc, 5:11], rep.emacro[:], rtol=1e-10, atol=1e-16)
        assert np.allclose(curve[inc, 11:17], rep.savg[:], rtol=1e-10, atol=1e-9)
    sig = np.fromfile(tmp_path / "stress.bin").reshape(6, 32, 32, 32)
    assert np.abs(sig - s.get_field(api.FIELD_STRESS)).max() < 1e-9 * np.abs(sig).max()
    assert curve[-1, 13] > curve[0, 13] > 0 and np.abs(curve[:, 11:13]).max() < 1e-2   # S33 rises, lateral stresses ~ 0
