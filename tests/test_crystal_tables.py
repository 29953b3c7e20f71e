"""CPU: the product library's slip / twin system tables (host_tables.cpp: closed-form Cartesian formulas + rotations)
against an independent construction from Miller(-Bravais) indices by lattice-vector arithmetic
(tests/golden/crystal_tables.py), plus physical sanity of every family.  The golden vectors are generated with the
independent tables, so the tables are also pinned through every golden parity test."""
import os
import sys

import numpy as np
import pytest

from lapx_b200 import microstructure as ms

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import crystal_tables as ct  # noqa: E402


def _tables(ph):
    ns = ph.nsys
    b = np.array([[ph.b[s][k] for k in range(3)] for s in range(ns)])
    n = np.array([[ph.n[s][k] for k in range(3)] for s in range(ns)])
    return b, n, np.array([ph.mode[s] for s in range(ns)])


def test_fcc_table_matches_miller_indices(product_lib):
    b, n, mode = _tables(ms.fcc_phase(product_lib))
    bi, ni = ct.fcc_systems()
    assert np.abs(b - bi).max() < 1e-15 and np.abs(n - ni).max() < 1e-15 and not mode.any()
    assert np.allclose(np.array(ms.fcc_phase(product_lib).c_voigt[:]).reshape(6, 6), ct.cubic_voigt(ms.CU_C11, ms.CU_C12, ms.CU_C44))
    # {111}<110>: every plane carries three directions, every direction lies in two planes
    assert np.abs(np.einsum("si,si->s", b, n)).max() < 1e-15


@pytest.mark.parametrize("with_twin", [0, 1, 2])
@pytest.mark.parametrize("covera", [ms.ZR_COVERA, 1.587, 1.624, 1.856])
def test_hcp_table_matches_miller_bravais_indices(covera, with_twin, product_lib):
    ph = ms.hcp_phase(product_lib, covera=covera, with_twin=with_twin)
    b, n, mode = _tables(ph)
    bi, ni, mi = ct.hcp_systems(covera, with_twin)
    assert np.array_equal(mode, mi)
    assert np.abs(b - bi).max() < 1e-14 and np.abs(n - ni).max() < 1e-14       # order and sign included
    assert np.allclose(np.array(ph.c_voigt[:]).reshape(6, 6), ct.hex_voigt(*ms.ZR_C5))
    for m in range(3, 3 + with_twin):
        assert ph.twin[m] == 1 and abs(ph.twin_shear[m] - ct.twin_shear(covera, m)) < 1e-14
    # physics of the families
    assert np.abs(np.einsum("si,si->s", b, n)).max() < 1e-14                  # b in the plane
    assert np.abs(n[mode == 0, 2]).max() < 1e-14 and np.abs(b[mode == 0, 2]).max() < 1e-14   # prismatic <a>: all in the basal plane
    assert np.allclose(np.abs(n[mode == 1, 2]), 1.0) and np.abs(b[mode == 1, 2]).max() < 1e-14   # basal <a>
    ca_len = np.sqrt(1.0 + covera**2)
    assert np.allclose(b[mode == 2, 2], covera / ca_len)                      # <c+a> = a + c: c component c/|c+a|
    if with_twin >= 1:
        # {10-12} tensile twin: a positive shear EXTENDS the c axis for c/a < sqrt 3 (Zr, Ti, Mg) and shortens it above
        e33 = b[mode == 3, 2] * n[mode == 3, 2]
        assert np.all(e33 > 0) if covera < np.sqrt(3.0) else np.all(e33 < 0)
    if with_twin >= 2:
        assert np.all(b[mode == 4, 2] * n[mode == 4, 2] < 0)                  # {11-22} compressive twin: contraction along c
    # six-fold symmetry: the set of Schmid tensors of a mode is invariant under a 60 degree rotation about c
    c6 = np.array([[0.5, -np.sqrt(3) / 2, 0], [np.sqrt(3) / 2, 0.5, 0], [0, 0, 1.0]])
    for m in np.unique(mode):
        S = 0.5 * (np.einsum("si,sj->sij", b[mode == m], n[mode == m]) + np.einsum("si,sj->sij", n[mode == m], b[mode == m]))
        Sr = np.einsum("ia,sab,jb->sij", c6, S, c6)
        twin = m >= 3
        for t in Sr:
            d = np.abs(S - t).reshape(len(S), -1).max(axis=1)
            if not twin:
                d = np.minimum(d, np.abs(S + t).reshape(len(S), -1).max(axis=1))   # slip: +-b equivalent
            assert d.min() < 1e-13
