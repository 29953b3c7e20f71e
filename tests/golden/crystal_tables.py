"""Slip / twin system tables built in numpy from Miller(-Bravais) INDICES — independent of the product library's
host_tables.cpp (which uses closed-form Cartesian formulas + 60-degree rotations).

Used by gen_golden.py for the golden vectors and by tests/test_crystal_tables.py to check the product tables entry by
entry (order and sign included: per-system fields such as the CRSS are indexed by system).  The index lists are the
standard families of the VPSC / EVPFFT literature:
    FCC  {111}<110>                                     12 systems
    HCP  prismatic <a>   {10-10}<-12-10>                 3
         basal <a>       (0001)<2-1-10>                  3
         pyramidal <c+a> {10-11}<-1-123>                12
         tensile twin    {10-12}<-1011>                  6   (extension along c for c/a < sqrt 3; direction reversed above)
         compressive twin {11-22}<11-2-3>                6   (contraction along c)
Construction: direct lattice a1, a2, a3 = -(a1+a2), c; a direction [uvtw] is u a1 + v a2 + t a3 + w c; a plane (hkil) has
the normal h a1* + k a2* + l c* of the three-axis reciprocal lattice (a1*, a2*, c*) of (a1, a2, c).
"""
import numpy as np

FCC_PLANES = [(1, 1, 1)] * 3 + [(-1, 1, 1)] * 3 + [(1, -1, 1)] * 3 + [(1, 1, -1)] * 3
FCC_DIRS = [(0, 1, -1), (1, 0, -1), (1, -1, 0), (0, 1, -1), (1, 0, 1), (1, 1, 0),
            (0, 1, 1), (1, 0, -1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, -1, 0)]

# (mode, plane hkil, direction uvtw)
HCP_SYSTEMS = [
    (0, (1, 0, -1, 0), (-1, 2, -1, 0)), (0, (0, 1, -1, 0), (-2, 1, 1, 0)), (0, (-1, 1, 0, 0), (-1, -1, 2, 0)),
    (1, (0, 0, 0, 1), (2, -1, -1, 0)), (1, (0, 0, 0, 1), (1, 1, -2, 0)), (1, (0, 0, 0, 1), (-1, 2, -1, 0)),
    (2, (1, 0, -1, 1), (-1, -1, 2, 3)), (2, (0, 1, -1, 1), (1, -2, 1, 3)), (2, (-1, 1, 0, 1), (2, -1, -1, 3)),
    (2, (-1, 0, 1, 1), (1, 1, -2, 3)), (2, (0, -1, 1, 1), (-1, 2, -1, 3)), (2, (1, -1, 0, 1), (-2, 1, 1, 3)),
    (2, (1, 0, -1, 1), (-2, 1, 1, 3)), (2, (0, 1, -1, 1), (-1, -1, 2, 3)), (2, (-1, 1, 0, 1), (1, -2, 1, 3)),
    (2, (-1, 0, 1, 1), (2, -1, -1, 3)), (2, (0, -1, 1, 1), (1, 1, -2, 3)), (2, (1, -1, 0, 1), (-1, 2, -1, 3)),
    (3, (1, 0, -1, 2), (-1, 0, 1, 1)), (3, (0, 1, -1, 2), (0, -1, 1, 1)), (3, (-1, 1, 0, 2), (1, -1, 0, 1)),
    (3, (-1, 0, 1, 2), (1, 0, -1, 1)), (3, (0, -1, 1, 2), (0, 1, -1, 1)), (3, (1, -1, 0, 2), (-1, 1, 0, 1)),
    (4, (1, 1, -2, 2), (1, 1, -2, -3)), (4, (-1, 2, -1, 2), (-1, 2, -1, -3)), (4, (-2, 1, 1, 2), (-2, 1, 1, -3)),
    (4, (-1, -1, 2, 2), (-1, -1, 2, -3)), (4, (1, -2, 1, 2), (1, -2, 1, -3)), (4, (2, -1, -1, 2), (2, -1, -1, -3)),
]


def _unit(v):
    v = np.asarray(v, float)
    return v / np.linalg.norm(v)


def fcc_systems():
    """(b, n): unit slip directions and plane normals, shape (12, 3) each."""
    b = np.array([_unit(d) for d in FCC_DIRS])
    n = np.array([_unit(p) for p in FCC_PLANES])
    return b, n


def hcp_lattice(covera):
    a1 = np.array([1.0, 0.0, 0.0])
    a2 = np.array([-0.5, np.sqrt(3.0) / 2.0, 0.0])
    c = np.array([0.0, 0.0, covera])
    vol = a1 @ np.cross(a2, c)
    rec = (np.cross(a2, c) / vol, np.cross(c, a1) / vol, np.cross(a1, a2) / vol)     # a1*, a2*, c*
    return (a1, a2, -(a1 + a2), c), rec


def hcp_systems(covera, with_twin=1):
    """(b, n, mode) for 18 / 24 / 30 systems (with_twin = 0 / 1 / 2)."""
    (a1, a2, a3, c), (r1, r2, rc) = hcp_lattice(covera)
    nmodes = 3 + with_twin
    b, n, mode = [], [], []
    for m, (h, k, i, l), (u, v, t, w) in HCP_SYSTEMS:
        if m >= nmodes:
            continue
        assert h + k + i == 0 and u + v + t == 0
        d = _unit(u * a1 + v * a2 + t * a3 + w * c)
        if m == 3 and covera**2 > 3.0:
            d = -d          # the {10-12} twinning shear reverses for c/a > sqrt 3 (Zn, Cd): contraction of c
        b.append(d)
        n.append(_unit(h * r1 + k * r2 + l * rc))
        mode.append(m)
    return np.array(b), np.array(n), np.array(mode)


def twin_shear(covera, m):
    """Characteristic shear of the {10-12} (m = 3) and {11-22} (m = 4) twins (Yoo 1981)."""
    if m == 3:
        return abs(covera**2 - 3.0) / (np.sqrt(3.0) * covera)
    return 2.0 * (covera**2 - 2.0) / (3.0 * covera)


def cubic_voigt(c11, c12, c44):
    C = np.zeros((6, 6))
    C[:3, :3] = c12
    C[np.arange(3), np.arange(3)] = c11
    C[np.arange(3, 6), np.arange(3, 6)] = c44
    return C


def hex_voigt(c11, c12, c13, c33, c44):
    C = np.zeros((6, 6))
    C[0, 0] = C[1, 1] = c11
    C[0, 1] = C[1, 0] = c12
    C[0, 2] = C[2, 0] = C[1, 2] = C[2, 1] = c13
    C[2, 2] = c33
    C[3, 3] = C[4, 4] = c44
    C[5, 5] = 0.5 * (c11 - c12)
    return C


def fill_phase(ph, b, n, mode, c_voigt, twin_modes=(), gamma0=1.0, nrate=10.0, tau0=(), voce=None, shears=None,
               thr=(0.1, 0.5)):
    """Fill an api.Phase (ctypes) from numpy tables; nothing comes from the product library."""
    ns = len(b)
    nm = int(max(mode)) + 1 if ns else 0
    ph.nsys, ph.nmodes = ns, nm
    for k, v in enumerate(np.asarray(c_voigt, float).reshape(36)):
        ph.c_voigt[k] = float(v)
    for s in range(ns):
        for k in range(3):
            ph.b[s][k] = float(b[s][k])
            ph.n[s][k] = float(n[s][k])
        ph.mode[s] = int(mode[s])
    for m in range(nm):
        ph.twin[m] = 1 if m in twin_modes else 0
        ph.gamma0[m] = gamma0
        ph.nrate[m] = nrate
        ph.tau0[m] = float(tau0[min(m, len(tau0) - 1)])
        if voce is not None:
            vm = voce[min(m, len(voce) - 1)]
            ph.tau1[m], ph.theta0[m], ph.theta1[m] = float(vm[0]), float(vm[1]), float(vm[2])
        for m2 in range(nm):
            ph.hlat[m][m2] = 1.0
        if shears and m in shears:
            ph.twin_shear[m] = float(shears[m])
    ph.twin_thr1, ph.twin_thr2 = thr
    return ph
