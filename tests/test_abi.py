"""CPU: the C-ABI libraries load and export every symbol include/evpfft.h declares; the product
library refuses to run without a GPU (no CPU fallback); argument errors behave as documented."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from lapx_b200 import api, microstructure as ms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "evpfft.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(evp_[a-z0-9_]+)\s*\(", txt)))


def test_header_lists_match_binding():
    syms = header_symbols()
    assert set(syms) == set(api.ABI_SYMBOLS_COMMON + api.ABI_SYMBOLS_PRODUCT_ONLY)


def test_product_exports_every_symbol(product_lib):
    for s in header_symbols():
        assert hasattr(product_lib, s), s
    assert product_lib.evp_abi_version() == 2
    assert product_lib.evp_backend() == b"cuda-sm100a"


def test_oracle_exports_common_symbols(oracle_lib):
    for s in api.ABI_SYMBOLS_COMMON:
        assert hasattr(oracle_lib, s), s
    assert oracle_lib.evp_backend() == b"cpu-oracle"
    assert oracle_lib.evp_abi_version() == 2


def test_build_id_is_the_digest_of_the_sources_in_the_tree(product_lib):
    """build() decides staleness by content: the digest embedded in the binary equals the digest of the sources next
    to it (a stale prebuilt .so would fail here and be rebuilt by build_product())."""
    from lapx_b200 import build
    assert product_lib.evp_build_id().decode() == "EVPSRC:" + build.source_id()


def test_struct_sizes_match_header():
    """ctypes mirrors of the ABI structs: sizes computed from the header's field lists (LP64)."""
    assert C.sizeof(api.Dist) == 4 * 4 + 128 + 4 + 12
    assert C.sizeof(api.IterReport) == 4 + 4 + 8 * 3 + 48 + 48 + 4 + 4 + 8
    assert C.sizeof(api.Ctrl) == 8 + 8 + 4 + 4 + 8 + 4 + 4 + 4 + 4   # trailing pad to 8
    assert C.sizeof(api.StepReport) == 4 + 4 + 16 + 48 * 3 + 8 * 3 + 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_gpu(product_lib):
    ph = ms.fcc_phase(product_lib)
    with pytest.raises(api.EvpError) as ei:
        api.Solver(product_lib, (8, 8, 8), [ph])
    assert ei.value.code == -3  # EVP_ERR_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_argument_errors(oracle_lib, product_lib):
    ph = ms.fcc_phase(product_lib)
    with pytest.raises(api.EvpError):
        api.Solver(oracle_lib, (1, 8, 8), [ph])
    s = api.Solver(oracle_lib, (8, 8, 8), [ph])
    with pytest.raises(api.EvpError) as ei:
        s.begin_increment(1e-3)          # nothing set yet
    assert ei.value.code == -2           # EVP_ERR_STATE
    ids, rot = ms.voronoi(product_lib, (8, 8, 8), 3)
    s.set_microstructure(ids, None, ms.expand_rotations(ids, rot))
    s.set_reference_medium(None)
    bad = api.Loading.uniaxial_tension(1.0)
    bad.iscau[2] = 1                     # both strain rate and stress imposed on 33
    with pytest.raises(api.EvpError) as ei:
        s.set_loading(bad)
    assert ei.value.code == -1
    with pytest.raises(api.EvpError):
        s.set_microstructure(ids, np.full(ids.shape, 3, np.int32), ms.expand_rotations(ids, rot))
    with pytest.raises(api.EvpError):
        s.op_green()                     # outside an increment


def test_crystal_tables(product_lib):
    """Schmid tables: n.b = 0, unit vectors, counts per mode (FCC 12; HCP 3+3+12+6(+6))."""
    for ph, counts in [(ms.fcc_phase(product_lib), [12]),
                       (ms.hcp_phase(product_lib, with_twin=0), [3, 3, 12]),
                       (ms.hcp_phase(product_lib, with_twin=1), [3, 3, 12, 6]),
                       (ms.hcp_phase(product_lib, with_twin=2), [3, 3, 12, 6, 6])]:
        ns = ph.nsys
        assert ns == sum(counts) and ph.nmodes == len(counts)
        b = np.array([[ph.b[s][k] for k in range(3)] for s in range(ns)])
        n = np.array([[ph.n[s][k] for k in range(3)] for s in range(ns)])
        assert np.allclose(np.linalg.norm(b, axis=1), 1) and np.allclose(np.linalg.norm(n, axis=1), 1)
        assert np.abs(np.einsum("si,si->s", b, n)).max() < 1e-14
        modes = [ph.mode[s] for s in range(ns)]
        assert [modes.count(m) for m in range(len(counts))] == counts
        # all systems distinct (up to sign)
        m = 0.5 * (np.einsum("si,sj->sij", b, n) + np.einsum("si,sj->sij", n, b)).reshape(ns, 9)
        for i in range(ns):
            for j in range(i + 1, ns):
                assert min(np.abs(m[i] - m[j]).max(), np.abs(m[i] + m[j]).max()) > 1e-6
    zr = ms.hcp_phase(product_lib, with_twin=1)
    assert abs(zr.twin_shear[3] - abs(ms.ZR_COVERA**2 - 3) / (np.sqrt(3) * ms.ZR_COVERA)) < 1e-14


def test_voronoi_bit_exact_vs_numpy(product_lib):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from voronoi_ref import voronoi_ids
    for grid, ng, seed in [((8, 8, 8), 5, 0), ((16, 12, 10), 50, 1), ((32, 32, 32), 50, 0), ((20, 20, 20), 700, 9)]:
        ids, rot = ms.voronoi(product_lib, grid, ng, seed)
        assert np.array_equal(ids, voronoi_ids(*grid, ng, seed))
        assert np.abs(np.einsum("gij,gkj->gik", rot, rot) - np.eye(3)).max() < 1e-14
    # slab generation equals the corresponding planes of the full tessellation
    full, _ = ms.voronoi(product_lib, (16, 16, 16), 40, 2)
    part, _ = ms.voronoi(product_lib, (16, 16, 16), 40, 2, z0=4, nzl=8)
    assert np.array_equal(part, full[4:12])
