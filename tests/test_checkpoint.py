"""Checkpoint / restart: the raw state file is the same for both back ends; a restarted run
continues bit-identically (same back end) and a checkpoint written by one back end restarts the other."""
import numpy as np
import pytest

from common import make_polycrystal, rel_err
from lapx_b200 import api


def _prepare(lib, host, texture=1):
    s, ids, grot = make_polycrystal(lib, host, (16, 16, 16), 10, seed=6)
    s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=6, itmin=1, tol_newton=1e-9, newton_itmax=100, update_texture=texture)
    s.set_loading(api.Loading.uniaxial_tension(1.0))
    return s


def _roundtrip(lib_a, lib_b, host, tmp_path):
    a = _prepare(lib_a, host)
    a.step(2e-4)
    a.step(2e-4)
    path = tmp_path / "state.ckpt"
    a.save_state(path)
    ra = a.step(2e-4)
    b = _prepare(lib_b, host)
    b.load_state(path)
    rb = b.step(2e-4)
    return a, b, ra, rb


def test_oracle_checkpoint_roundtrip(oracle_lib, product_lib, tmp_path):
    a, b, ra, rb = _roundtrip(oracle_lib, oracle_lib, product_lib, tmp_path)
    # (the oracle's OpenMP reductions are not order-deterministic, hence a rounding-level tolerance instead of bit equality)
    assert rel_err(a.get_field(api.FIELD_STRESS), b.get_field(api.FIELD_STRESS)) < 1e-12
    assert rel_err(a.get_field(api.FIELD_ROTATION), b.get_field(api.FIELD_ROTATION)) < 1e-12
    assert rel_err(ra.savg[:], rb.savg[:]) < 1e-12 and rel_err(ra.emacro[:], rb.emacro[:]) < 1e-12
    bad = api.Solver(oracle_lib, (8, 8, 8), [_phase(product_lib)])
    with pytest.raises(api.EvpError):
        bad.load_state(tmp_path / "state.ckpt")


def _phase(host):
    from lapx_b200 import microstructure as ms
    return ms.fcc_phase(host)


@pytest.mark.gpu
def test_gpu_checkpoint_roundtrip_and_cross_backend(oracle_lib, product_lib, tmp_path):
    a, b, ra, rb = _roundtrip(product_lib, product_lib, product_lib, tmp_path)
    # the CUDA path is deterministic (fixed-order reductions on device and host): the restart is bit identical
    assert np.array_equal(a.get_field(api.FIELD_STRESS), b.get_field(api.FIELD_STRESS))
    assert ra.savg[:] == rb.savg[:]
    # a checkpoint written by the CUDA path restarts the oracle (and the continued runs agree to 1e-8)
    c = _prepare(oracle_lib, product_lib)
    c.load_state(tmp_path / "state.ckpt")
    rc = c.step(2e-4)
    assert rel_err(rc.savg[:], ra.savg[:]) < 1e-8
    assert rel_err(c.get_field(api.FIELD_STRESS), a.get_field(api.FIELD_STRESS)) < 1e-8
