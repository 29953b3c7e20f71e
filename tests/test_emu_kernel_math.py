"""CPU: the thread-level math of the CUDA kernels (lapx_b200/csrc/evp_core.h) executed on the host
through tests/emu (same inline functions the __global__ kernels call) against numpy references."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from lapx_b200 import api, microstructure as ms

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "emu.cpp")
    out = os.path.join(HERE, "emu", "libevp_emu.so")
    deps = [src] + [os.path.join(HERE, "..", "lapx_b200", "csrc", f) for f in ("evp_core.h", "host_math.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", out, src], check=True)
    return C.CDLL(out)


@pytest.mark.parametrize("n", [8, 16, 32, 64, 128, 256, 512, 1024])
@pytest.mark.parametrize("inv", [0, 1])
def test_stockham_passes_match_numpy(emu, n, inv):
    rng = np.random.default_rng(n + inv)
    x = rng.normal(size=(3, n)) + 1j * rng.normal(size=(3, n))
    buf = np.ascontiguousarray(x.copy())
    assert emu.emu_fft(n, inv, 3, buf.ctypes.data_as(C.c_void_p)) == 0
    ref = np.fft.ifft(x, axis=1) * n if inv else np.fft.fft(x, axis=1)
    assert np.abs(buf - ref).max() < 5e-13 * np.abs(ref).max()


@pytest.mark.parametrize("n", [128, 256, 512])
@pytest.mark.parametrize("inv", [0, 1])
def test_radix16_passes_match_numpy(emu, n, inv):
    rng = np.random.default_rng(3 * n + inv)
    x = rng.normal(size=(3, n)) + 1j * rng.normal(size=(3, n))
    buf = np.ascontiguousarray(x.copy())
    assert emu.emu_fft16(n, inv, 3, buf.ctypes.data_as(C.c_void_p)) == 0
    ref = np.fft.ifft(x, axis=1) * n if inv else np.fft.fft(x, axis=1)
    assert np.abs(buf - ref).max() < 5e-13 * np.abs(ref).max()


def _rand_rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


PAIRS = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]
VM = np.array([[0, 5, 4], [5, 1, 3], [4, 3, 2]])


def _c4(cv):
    C4 = np.zeros((3, 3, 3, 3))
    for i in range(3):
        for j in range(3):
            for k in range(3):
                for l in range(3):
                    C4[i, j, k, l] = cv[VM[i, j], VM[k, l]]
    return C4


def _sym(v):
    A = np.zeros(v.shape[:-1] + (3, 3), dtype=v.dtype)
    for a, (i, j) in enumerate(PAIRS):
        A[..., i, j] = v[..., a]
        A[..., j, i] = v[..., a]
    return A


def _c6(A):
    return np.stack([A[..., i, j] for (i, j) in PAIRS], axis=-1)


def _aniso_c0(rng, product_lib):
    ph = ms.fcc_phase(product_lib)
    cv = np.array(list(ph.c_voigt)).reshape(6, 6)
    C = _c4(cv)
    acc = np.zeros((3, 3, 3, 3))
    for _ in range(3):
        R = _rand_rot(rng)
        acc += np.einsum("ia,jb,kc,ld,abcd->ijkl", R, R, R, R, C)
    acc /= 3
    out = np.zeros((6, 6))
    for a, (i, j) in enumerate(PAIRS):
        for b, (k, l) in enumerate(PAIRS):
            out[a, b] = acc[i, j, k, l]
    return 0.5 * (out + out.T)


def test_green_point_matches_tensor_formula(emu, product_lib):
    rng = np.random.default_rng(5)
    c0 = _aniso_c0(rng, product_lib)
    C4 = _c4(c0)
    emu.emu_green_point.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double,
                                    C.c_void_p, C.c_void_p]
    for trial in range(20):
        xi = rng.normal(size=3)
        lam6 = rng.normal(size=6) + 1j * rng.normal(size=6)
        out = np.zeros(6, complex)
        lam_c = np.ascontiguousarray(lam6)
        emu.emu_green_point(c0.ctypes.data_as(C.c_void_p), xi[0], xi[1], xi[2], 0, 0, 0.25,
                            lam_c.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        A = np.einsum("ijkl,j,l->ik", C4, xi, xi)
        G = np.linalg.inv(A)
        gam = 0.25 * (np.einsum("ik,j,l->ijkl", G, xi, xi) + np.einsum("jk,i,l->ijkl", G, xi, xi)
                      + np.einsum("il,j,k->ijkl", G, xi, xi) + np.einsum("jl,i,k->ijkl", G, xi, xi))
        ref = _c6(np.einsum("ijkl,kl->ij", gam, _sym(lam6))) * 0.25
        assert np.abs(out - ref).max() < 1e-12 * np.abs(ref).max()
        # Nyquist rule: S0 : lam
        emu.emu_green_point(c0.ctypes.data_as(C.c_void_p), xi[0], xi[1], xi[2], 0, 1, 1.0,
                            lam_c.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        W = np.array([1, 1, 1, np.sqrt(2), np.sqrt(2), np.sqrt(2)])
        S0m = np.linalg.inv(c0 * np.outer(W, W))
        ref = (S0m @ (W * lam6)) / W
        assert np.abs(out - ref).max() < 1e-12 * np.abs(ref).max()
    emu.emu_green_point(c0.ctypes.data_as(C.c_void_p), 0.0, 0.0, 0.0, 1, 0, 1.0, lam_c.ctypes.data_as(C.c_void_p),
                        out.ctypes.data_as(C.c_void_p))
    assert np.all(out == 0)


@pytest.mark.parametrize("hcp,nrate,variants", [(False, 10.0, [0, 1, 2, 11, 12, 16]), (True, 10.0, [0, 3, 6, 13, 18]), (False, 20.0, [0, 2, 4, 14, 17]),
                                                 (True, 20.0, [5, 15, 19]), (False, 7.5, [0, 2]), (2, 10.0, [0, 20])])
@pytest.mark.parametrize("iso", [False, True])
def test_constitutive_voxel_matches_sample_frame_newton(emu, product_lib, hcp, nrate, variants, iso):
    """Crystal-frame b-basis LDL^T Newton (kernel math) vs a sample-frame Mandel Newton in numpy."""
    rng = np.random.default_rng(11 + hcp + 2 * iso + int(nrate))
    # hcp: True = 24 systems (tensile twins), 2 = 30 systems (+ compressive twins: the 30-system fast path, variant 20)
    ph = ms.hcp_phase(product_lib, with_twin=int(hcp), nrate=nrate) if hcp else ms.fcc_phase(product_lib, nrate=nrate)
    if iso:
        K, mu = 140000.0, 48000.0
        c0 = np.zeros((6, 6))
        c0[:3, :3] = K - 2 * mu / 3
        c0[np.arange(3), np.arange(3)] = K + 4 * mu / 3
        c0[np.arange(3, 6), np.arange(3, 6)] = mu
    else:
        c0 = _aniso_c0(rng, product_lib)
    W = np.array([1, 1, 1, np.sqrt(2), np.sqrt(2), np.sqrt(2)])
    C0m = c0 * np.outer(W, W)
    S0m = np.linalg.inv(C0m)
    ns = ph.nsys
    b = np.array([[ph.b[s][k] for k in range(3)] for s in range(ns)])
    n = np.array([[ph.n[s][k] for k in range(3)] for s in range(ns)])
    mode = np.array([ph.mode[s] for s in range(ns)])
    twin = np.array([ph.twin[m] for m in mode]).astype(bool)
    nr = np.array([ph.nrate[m] for m in mode])
    g0 = np.array([ph.gamma0[m] for m in mode])
    cv = np.array(list(ph.c_voigt)).reshape(6, 6)
    emu.emu_constitutive.argtypes = [C.c_void_p] * 7 + [C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 3 + [C.c_int]
    emu.emu_constitutive_t.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 3
    for trial in range(25):
        R = _rand_rot(rng)
        Cs = np.einsum("ia,jb,kc,ld,abcd->ijkl", R, R, R, R, _c4(cv))
        Sm = np.linalg.inv(np.array([[Cs[i, j, k, l] * W[a] * W[bb] for bb, (k, l) in enumerate(PAIRS)]
                                     for a, (i, j) in enumerate(PAIRS)]))
        m = 0.5 * (np.einsum("si,sj->sij", b, n) + np.einsum("si,sj->sij", n, b))
        ms6 = _c6(np.einsum("ia,sab,jb->sij", R, m, R)) * W
        crss = rng.uniform(10, 40, size=ns)
        so = rng.normal(size=6) * 15.0
        e = rng.normal(size=6) * 3e-4
        ep = rng.normal(size=6) * 1e-4
        dt = 2e-4
        so_m, e_m, ep_m = W * so, W * e, W * ep
        s = so_m.copy()
        for it in range(200):
            tau = ms6 @ s
            x = np.abs(tau) / crss
            gd = g0 * x ** nr * np.sign(tau)
            dgd = g0 * nr * x ** (nr - 1) / crss
            off = twin & (tau <= 0)
            gd[off] = 0
            dgd[off] = 0
            F = S0m @ (s - so_m) + Sm @ s + ep_m + dt * (gd @ ms6) - e_m
            J = S0m + Sm + dt * np.einsum("s,sa,sb->ab", dgd, ms6, ms6)
            d = np.linalg.solve(J, -F)
            s = s + d
            if np.linalg.norm(d) <= 1e-13 * np.linalg.norm(s):
                break
        sig = np.ascontiguousarray(so.copy())
        ds, de, bad = C.c_double(), C.c_double(), C.c_int()
        Rc = np.ascontiguousarray(R)
        nit = emu.emu_constitutive(C.byref(ph), c0.ctypes.data_as(C.c_void_p), Rc.ctypes.data_as(C.c_void_p),
                                   sig.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                                   ep.ctypes.data_as(C.c_void_p), crss.ctypes.data_as(C.c_void_p), dt, 1e-12, 200,
                                   C.byref(ds), C.byref(de), C.byref(bad), 0)
        assert bad.value == 0 and nit < 200
        ref = s / W
        assert np.abs(sig - ref).max() < 1e-9 * np.abs(ref).max(), (trial, sig, ref)
        assert abs(ds.value - np.linalg.norm(s - so_m)) < 1e-9 * np.linalg.norm(s - so_m)
        assert abs(de.value - np.linalg.norm(S0m @ (s - so_m))) < 1e-9 * np.linalg.norm(S0m @ (s - so_m))
        for var in variants:   # production (templated / unrolled) kernel variants
            sig3 = np.ascontiguousarray(so.copy())
            nit3 = emu.emu_constitutive_t(var, C.byref(ph), c0.ctypes.data_as(C.c_void_p), Rc.ctypes.data_as(C.c_void_p),
                                          sig3.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                                          ep.ctypes.data_as(C.c_void_p), crss.ctypes.data_as(C.c_void_p), dt, 1e-12, 200,
                                          C.byref(ds), C.byref(de), C.byref(bad))
            assert bad.value == 0 and nit3 == nit, (var, nit3, nit)
            assert np.abs(sig3 - ref).max() < 1e-9 * np.abs(ref).max(), var
            assert abs(ds.value - np.linalg.norm(s - so_m)) < 1e-9 * np.linalg.norm(s - so_m)
            assert abs(de.value - np.linalg.norm(S0m @ (s - so_m))) < 1e-9 * np.linalg.norm(S0m @ (s - so_m))
        if iso:  # the general (rotated S0) path must agree with the isotropic fast path
            sig2 = np.ascontiguousarray(so.copy())
            emu.emu_constitutive(C.byref(ph), c0.ctypes.data_as(C.c_void_p), Rc.ctypes.data_as(C.c_void_p),
                                 sig2.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                                 ep.ctypes.data_as(C.c_void_p), crss.ctypes.data_as(C.c_void_p), dt, 1e-12, 200,
                                 C.byref(ds), C.byref(de), C.byref(bad), 1)
            assert np.abs(sig2 - sig).max() < 1e-11 * np.abs(sig).max()
