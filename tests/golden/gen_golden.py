"""Generate tests/golden/*.npz — golden vectors for the EVPFFT equilibrium loop.

PARITY UNPINNED: /root/reference holds only LICENSE, so these vectors do NOT come from LApx.
They come from an independent numpy/scipy restatement of the published algorithm (SURVEY.md
§8(a)) written in this file: full 3x3x3x3 tensors + einsum, scipy.fft.rfftn, numpy.linalg — no
code shared with oracle/evp_oracle.cpp or the CUDA path.  The C++ oracle and the CUDA library
are both checked against the files this script writes (tests/test_oracle_golden.py,
tests/test_gpu_parity.py).

Run (CPU only, a few seconds):   python tests/golden/gen_golden.py
Needs the built product library only for the Voronoi grain ids (an input, itself checked bit-exactly against
oracle/voronoi_ref.py).  The slip / twin system tables are NOT taken from the product library: they are built here
from Miller(-Bravais) indices (tests/golden/crystal_tables.py), so a wrong table in host_tables.cpp shows up as a
golden mismatch.
"""
import os
import sys

import numpy as np
import scipy.fft as sfft

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lapx_b200 import api, microstructure as ms  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import crystal_tables as ct  # noqa: E402

VOCE_HCP = [[5.0, 100.0, 5.0], [10.0, 200.0, 10.0], [20.0, 400.0, 20.0], [5.0, 50.0, 5.0]]


def fcc_phase_np(gamma0=1.0, nrate=10.0, tau0=16.0, tau1=0.0, theta0=0.0, theta1=0.0):
    b, n = ct.fcc_systems()
    return ct.fill_phase(api.Phase(), b, n, np.zeros(12, int), ct.cubic_voigt(ms.CU_C11, ms.CU_C12, ms.CU_C44), gamma0=gamma0,
                         nrate=nrate, tau0=[tau0], voce=[[tau1, theta0, theta1]])


def hcp_phase_np(with_twin=1, nrate=10.0, tau0_mode=(20.0, 100.0, 160.0, 80.0), voce_mode=None, thr=(0.1, 0.5)):
    b, n, mode = ct.hcp_systems(ms.ZR_COVERA, with_twin)
    twin_modes = tuple(range(3, 3 + with_twin))
    return ct.fill_phase(api.Phase(), b, n, mode, ct.hex_voigt(*ms.ZR_C5), twin_modes=twin_modes, nrate=nrate, tau0=tau0_mode,
                         voce=voce_mode, shears={m: ct.twin_shear(ms.ZR_COVERA, m) for m in twin_modes}, thr=thr)


VM = np.array([[0, 5, 4], [5, 1, 3], [4, 3, 2]])
PAIRS = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]


def mandel_basis():
    B = np.zeros((6, 3, 3))
    for a, (i, j) in enumerate(PAIRS):
        if i == j:
            B[a, i, j] = 1.0
        else:
            B[a, i, j] = B[a, j, i] = 1.0 / np.sqrt(2.0)
    return B


B6 = mandel_basis()


def voigt_to_t4(cv):
    cv = np.asarray(cv).reshape(6, 6)
    C = np.zeros((3, 3, 3, 3))
    for i in range(3):
        for j in range(3):
            for k in range(3):
                for l in range(3):
                    C[i, j, k, l] = cv[VM[i, j], VM[k, l]]
    return C


def t4_to_m6(C):
    return np.einsum("aij,ijkl,bkl->ab", B6, C, B6)


def m6_to_t4(M):
    return np.einsum("aij,ab,bkl->ijkl", B6, M, B6)


def sym_to_m6(A):        # (...,3,3) -> (...,6)
    return np.einsum("aij,...ij->...a", B6, A)


def m6_to_sym(v):        # (...,6) -> (...,3,3)
    return np.einsum("aij,...a->...ij", B6, v)


def cart6_to_sym(f):     # ABI layout (6, ...) -> (..., 3, 3)
    A = np.zeros(f.shape[1:] + (3, 3))
    for a, (i, j) in enumerate(PAIRS):
        A[..., i, j] = f[a]
        A[..., j, i] = f[a]
    return A


def sym_to_cart6(A):
    return np.stack([A[..., i, j] for (i, j) in PAIRS], axis=0)


def ipow(x, n):
    ni = int(n)
    if ni == n and 0 <= ni <= 64:
        r = np.ones_like(x)
        b = x.copy()
        k = ni
        while k:
            if k & 1:
                r = r * b
            b = b * b
            k >>= 1
        return r
    return np.power(x, n)


class NumpyEVP:
    """Straight restatement of SURVEY.md §8(a) rows a1..a7 with dense tensors."""

    def __init__(self, phase: api.Phase, grain, grain_rot, c0_voigt=None):
        self.shape = grain.shape                      # (nz, ny, nx)
        nz, ny, nx = self.shape
        self.N = grain.size
        R = grain_rot[grain.reshape(-1)].copy()       # (N,3,3) crystal -> sample
        self.R = R
        ns = phase.nsys
        self.ns = ns
        b = np.array([[phase.b[s][k] for k in range(3)] for s in range(ns)])
        n = np.array([[phase.n[s][k] for k in range(3)] for s in range(ns)])
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        self.bc, self.nc = b, n
        self.mc = 0.5 * (np.einsum("si,sj->sij", b, n) + np.einsum("si,sj->sij", n, b))
        self.qc = 0.5 * (np.einsum("si,sj->sij", b, n) - np.einsum("si,sj->sij", n, b))   # skew part, crystal frame
        self.Cc = voigt_to_t4(np.array(list(phase.c_voigt)))
        Cc = self.Cc
        # rotate per grain, then gather
        Cg = np.einsum("gia,gjb,gkc,gld,abcd->gijkl", grain_rot, grain_rot, grain_rot, grain_rot, Cc)
        Cm6 = np.einsum("aij,gijkl,bkl->gab", B6, Cg, B6)
        self.rebuild_from_R()
        if c0_voigt is None:
            counts = np.bincount(grain.reshape(-1), minlength=len(grain_rot)).astype(float)
            C0m = np.einsum("g,gab->ab", counts, Cm6) / counts.sum()
            C0m = 0.5 * (C0m + C0m.T)
        else:
            C0m = t4_to_m6(voigt_to_t4(c0_voigt))
        self.C0m = C0m
        self.S0m = np.linalg.inv(C0m)
        self.C0t = m6_to_t4(C0m)
        mode = np.array([phase.mode[s] for s in range(ns)])
        self.nrate = np.array([phase.nrate[m] for m in mode])
        self.g0 = np.array([phase.gamma0[m] for m in mode])
        self.twin = np.array([phase.twin[m] for m in mode]).astype(bool)
        self.tau0 = np.array([phase.tau0[m] for m in mode])
        self.mode = mode
        self.phase = phase
        self.crss = np.tile(self.tau0, (self.N, 1))
        self.gacc = np.zeros(self.N)
        self.sig = np.zeros((self.N, 6))      # Mandel
        self.e = np.zeros((self.N, 6))
        self.epsp = np.zeros((self.N, 6))
        self.Et = np.zeros(6)                 # Cartesian comps
        self.E = np.zeros(6)
        self.dEpend = np.zeros(6)
        self.Edot_prev = np.zeros(6)
        self.tol_newton, self.newton_itmax = 1e-9, 100
        self.update_texture = self.update_twinning = False
        self.wrot = np.zeros((self.N, 3))
        self.twinf = np.zeros((self.N, ns))
        self.twinned = np.zeros(self.N, bool)
        self.facc = 0.0                       # accumulated twin fraction: history sum, never reduced by a reorientation
        self.tshear = np.array([phase.twin_shear[m] for m in mode])
        self.W = np.array([1, 1, 1, np.sqrt(2), np.sqrt(2), np.sqrt(2)])

    def rebuild_from_R(self):
        """Orientation-dependent quantities of every voxel from its current rotation."""
        R = self.R
        self.m = np.einsum("via,sab,vjb->vsij", R, self.mc, R)      # (N,ns,3,3) sample frame
        self.m6 = sym_to_m6(self.m)                                 # (N,ns,6)
        Cv = np.einsum("via,vjb,vkc,vld,abcd->vijkl", R, R, R, R, self.Cc)
        self.S6 = np.linalg.inv(np.einsum("aij,vijkl,bkl->vab", B6, Cv, B6))   # (N,6,6)

    def local_rotation(self):
        """Axial (w32,w13,w21) of skew(grad u) of the compatible strain field e; zero at xi=0 and Nyquist."""
        nz, ny, nx = self.shape
        eh = sfft.rfftn(m6_to_sym(self.e).reshape(nz, ny, nx, 3, 3), axes=(0, 1, 2))
        fz = sfft.fftfreq(nz) * nz
        fy = sfft.fftfreq(ny) * ny
        fx = sfft.rfftfreq(nx) * nx
        xi = np.stack(np.meshgrid(fz / nz, fy / ny, fx / nx, indexing="ij")[::-1], axis=-1)
        x2 = (xi ** 2).sum(-1)
        x2[0, 0, 0] = 1.0
        t = np.einsum("...ik,...k->...i", eh, xi)
        wh = (np.einsum("...i,...j->...ij", t, xi) - np.einsum("...j,...i->...ij", t, xi)) / x2[..., None, None]
        nyq = np.zeros((nz, ny, nx // 2 + 1), bool)
        if nz % 2 == 0:
            nyq[nz // 2, :, :] = True
        if ny % 2 == 0:
            nyq[:, ny // 2, :] = True
        if nx % 2 == 0:
            nyq[:, :, nx // 2] = True
        wh[nyq] = 0.0
        wh[0, 0, 0] = 0.0
        w = sfft.irfftn(wh, s=(nz, ny, nx), axes=(0, 1, 2)).reshape(self.N, 3, 3)
        return np.stack([w[:, 2, 1], w[:, 0, 2], w[:, 1, 0]], axis=1)

    # ---- loading ----
    def set_loading(self, ld: api.Loading):
        self.udot = np.asarray(ld.udot, float).reshape(3, 3)
        iud = np.asarray(ld.iudot).reshape(3, 3)
        self.strain_ctl = np.array([bool(iud[i, j] and iud[j, i]) for (i, j) in PAIRS])
        self.scau = np.asarray(ld.scau, float)

    def begin_increment(self, dt):
        self.dt = dt
        D = 0.5 * (self.udot + self.udot.T)
        for c, (i, j) in enumerate(PAIRS):
            rate = D[i, j] if self.strain_ctl[c] else self.Edot_prev[c]
            self.dEpend[c] = dt * rate
            self.E[c] = self.Et[c] + self.dEpend[c]

    # ---- rows a1..a3 ----
    def op_green(self):
        nz, ny, nx = self.shape
        sig = m6_to_sym(self.sig).reshape(nz, ny, nx, 3, 3)
        sh = sfft.rfftn(sig, axes=(0, 1, 2))
        fz = sfft.fftfreq(nz) * nz
        fy = sfft.fftfreq(ny) * ny
        fx = sfft.rfftfreq(nx) * nx
        # numpy puts the Nyquist index at -n/2; direction only enters through even functions
        xi = np.stack(np.meshgrid(fz / nz, fy / ny, fx / nx, indexing="ij")[::-1], axis=-1)  # (..,3) = (x,y,z)
        A = np.einsum("ijkl,...j,...l->...ik", self.C0t, xi, xi)
        A[0, 0, 0] = np.eye(3)
        G = np.linalg.inv(A)
        t = np.einsum("...kl,...l->...k", sh, xi)
        u = np.einsum("...ik,...k->...i", G, t)
        de_h = 0.5 * (np.einsum("...i,...j->...ij", u, xi) + np.einsum("...j,...i->...ij", u, xi))
        # Nyquist planes: Gamma := S0
        nyq = np.zeros((nz, ny, nx // 2 + 1), bool)
        if nz % 2 == 0:
            nyq[nz // 2, :, :] = True
        if ny % 2 == 0:
            nyq[:, ny // 2, :] = True
        if nx % 2 == 0:
            nyq[:, :, nx // 2] = True
        s6 = sym_to_m6(sh[nyq])
        de_h[nyq] = m6_to_sym(s6 @ self.S0m.T)
        de_h[0, 0, 0] = 0.0
        de = sfft.irfftn(de_h, s=(nz, ny, nx), axes=(0, 1, 2))
        self.de = sym_to_m6(de.reshape(self.N, 3, 3))
        self.e = self.e - self.de + (self.W * self.dEpend)[None, :]
        self.dEpend[:] = 0.0

    # ---- rows a4..a6 ----
    def rates(self, s6):
        tau = np.einsum("vsa,va->vs", self.m6, s6)
        x = np.abs(tau) / self.crss
        xn1 = np.stack([ipow(x[:, s], self.nrate[s] - 1.0) for s in range(self.ns)], axis=1)
        gd = self.g0 * xn1 * x * np.sign(tau)
        dgd = self.g0 * self.nrate * xn1 / self.crss
        off = self.twin[None, :] & (tau <= 0)
        gd = np.where(off, 0.0, gd)
        dgd = np.where(off, 0.0, dgd)
        return gd, dgd

    def op_constitutive(self):
        so = self.sig.copy()
        s = so.copy()
        active = np.ones(self.N, bool)
        nit = np.zeros(self.N, int)
        for _ in range(self.newton_itmax):
            if not active.any():
                break
            gd, dgd = self.rates(s)
            edp = np.einsum("vs,vsa->va", gd, self.m6)
            dedp = np.einsum("vs,vsa,vsb->vab", dgd, self.m6, self.m6)
            F = (s - so) @ self.S0m.T + np.einsum("vab,vb->va", self.S6, s) + self.epsp + self.dt * edp - self.e
            J = self.S0m[None] + self.S6 + self.dt * dedp
            d = np.linalg.solve(J, -F[..., None])[..., 0]
            d[~active] = 0.0
            s = s + d
            nit[active] += 1
            done = np.linalg.norm(d, axis=1) <= self.tol_newton * np.linalg.norm(s, axis=1)
            active &= ~done
        gd, _ = self.rates(s)
        edp = np.einsum("vs,vsa->va", gd, self.m6)
        eps = np.einsum("vab,vb->va", self.S6, s) + self.epsp + self.dt * edp
        errs = np.linalg.norm(s - so, axis=1).mean()
        erre = np.linalg.norm(eps - self.e, axis=1).mean()
        self.sig = s
        self.edotp = edp
        savg = (s / self.W).mean(axis=0)
        self.savg = savg
        w2 = np.array([1, 1, 1, 2, 2, 2.0])
        sn = np.sqrt((w2 * savg**2).sum())
        en = np.sqrt((w2 * self.E**2).sum())
        err_s = errs / sn if sn > 0 else errs
        err_e = erre / en if en > 0 else erre
        # a7
        T = np.where(~self.strain_ctl)[0]
        self.dEpend[:] = 0.0
        if len(T):
            r = self.W[T] * (self.scau[T] - savg[T])
            dEm = np.linalg.solve(self.C0m[np.ix_(T, T)], r)
            self.dEpend[T] = dEm / self.W[T]
            self.E[T] += self.dEpend[T]
        return dict(err_stress=err_s, err_strain=err_e, savg=savg.copy(), emacro=self.E.copy(),
                    newton_max=int(nit.max()), newton_mean=float(nit.mean()))

    def voce(self, m, G):
        p = self.phase
        t0, t1, h0, h1 = p.tau0[m], p.tau1[m], p.theta0[m], p.theta1[m]
        if abs(t1) < 1e-300:
            return t0 + h1 * G
        return t0 + (t1 + h1 * G) * (1.0 - np.exp(-G * abs(h0 / t1)))

    def end_increment(self):
        gd, _ = self.rates(self.sig)
        edp = np.einsum("vs,vsa->va", gd, self.m6)
        self.epsp = self.epsp + self.dt * edp
        dg = np.abs(gd) * self.dt
        dG = dg.sum(axis=1)
        hl = np.array([[self.phase.hlat[self.mode[s]][self.mode[s2]] for s2 in range(self.ns)] for s in range(self.ns)])
        for s in range(self.ns):
            m = self.mode[s]
            dv = self.voce(m, self.gacc + dG) - self.voce(m, self.gacc)
            hs = dg @ hl[s]
            with np.errstate(invalid="ignore", divide="ignore"):
                dtau = np.where(dG > 0, dv * hs / dG, 0.0)
            self.crss[:, s] = self.crss[:, s] + dtau
        self.gacc = self.gacc + dG
        nre = 0
        if self.update_twinning:
            tw = self.twin & (self.tshear > 0)
            df = gd[:, tw] * self.dt / self.tshear[tw]
            self.twinf[:, tw] += df
            self.facc += df.sum() / self.N
            Facc = self.facc
            Feff = self.twinned.sum() / self.N
        if self.update_texture:
            wnew = self.local_rotation()
            W = 0.5 * (self.udot - self.udot.T)
            wapp = np.array([W[2, 1], W[0, 2], W[1, 0]])
            qs = np.einsum("vs,sij->vij", gd, self.qc)                    # plastic spin, crystal frame
            wpc = np.stack([qs[:, 2, 1], qs[:, 0, 2], qs[:, 1, 0]], axis=1)
            wps = np.einsum("vij,vj->vi", self.R, wpc)
            dw = self.dt * wapp[None, :] + (wnew - self.wrot) - self.dt * wps
            self.wrot = wnew
            th = np.linalg.norm(dw, axis=1)
            K = np.zeros((self.N, 3, 3))
            K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -dw[:, 2], dw[:, 1], dw[:, 2], -dw[:, 0], -dw[:, 1], dw[:, 0]
            with np.errstate(invalid="ignore", divide="ignore"):
                a = np.where(th < 1e-8, 1 - th ** 2 / 6, np.sin(th) / th)
                b = np.where(th < 1e-8, 0.5 - th ** 2 / 24, (1 - np.cos(th)) / th ** 2)
            Q = np.eye(3)[None] + a[:, None, None] * K + b[:, None, None] * (K @ K)
            self.R = Q @ self.R
        if self.update_twinning:
            thr = self.phase.twin_thr1 + (self.phase.twin_thr2 * Feff / Facc if Facc > 0 else 0.0)
            ftw = np.where(self.twin[None, :], self.twinf, -1.0)
            best = ftw.argmax(axis=1)
            fb = ftw.max(axis=1)
            hit = (~self.twinned) & (fb > thr) & (fb > 0)
            for v in np.where(hit)[0]:
                nvec = self.nc[best[v]]
                self.R[v] = self.R[v] @ (2 * np.outer(nvec, nvec) - np.eye(3))
                self.twinf[v] = 0.0
                self.twinned[v] = True
            nre = int(hit.sum())
            self.last_twin = (Facc, self.twinned.sum() / self.N, nre)
        if self.update_texture or nre:
            self.rebuild_from_R()
        self.E = self.E - self.dEpend
        self.dEpend[:] = 0.0
        self.Edot_prev = (self.E - self.Et) / self.dt
        self.Et = self.E.copy()

    def field(self, v6):      # Mandel (N,6) -> ABI (6,nz,ny,nx)
        return np.ascontiguousarray((v6 / self.W).T).reshape((6,) + self.shape)


def make_case(lib, name, grid, ngrains, seed, loading, dt, iters_per_inc, nincs, phase=None, hcp=False, slim=False, texture=False,
              twinning=False, phase_kind=0):
    nx, ny, nz = grid
    ids, grot = ms.voronoi(lib, grid, ngrains, seed)
    if phase is None:
        phase = fcc_phase_np(gamma0=1.0, nrate=10.0, tau0=16.0, tau1=10.0, theta0=200.0, theta1=10.0)
    S = NumpyEVP(phase, ids, grot)
    S.update_texture, S.update_twinning = texture, twinning
    S.set_loading(loading)
    reports = []
    fields = {}
    twin_hist = []
    for inc in range(nincs):
        S.begin_increment(dt)
        for it in range(iters_per_inc):
            S.op_green()
            if inc == 0 and it == 1:
                fields["e_after_green_inc0_it2"] = S.field(S.e)
                fields["de_inc0_it2"] = S.field(S.de)
            r = S.op_constitutive()
            reports.append([inc, it + 1, r["err_stress"], r["err_strain"], *r["savg"], *r["emacro"], r["newton_max"],
                            r["newton_mean"]])
            if inc == 0 and it == 1:
                fields["sig_inc0_it2"] = S.field(S.sig)
        S.end_increment()
        fields[f"sig_end_inc{inc}"] = S.field(S.sig)
        fields[f"e_end_inc{inc}"] = S.field(S.e)
        fields[f"epsp_end_inc{inc}"] = S.field(S.epsp)
        fields[f"crss_end_inc{inc}"] = np.ascontiguousarray(S.crss.T).reshape((S.ns, nz, ny, nx))
        if texture or twinning:
            fields[f"rot_end_inc{inc}"] = np.ascontiguousarray(S.R.reshape(S.N, 9).T).reshape((9, nz, ny, nx))
            fields[f"twinned_end_inc{inc}"] = S.twinned.reshape(nz, ny, nx).astype(np.int32)
            if twinning:
                print("   inc", inc, "twin F_acc, F_eff, reoriented:", S.last_twin)
                twin_hist.append(list(S.last_twin))
    if slim:  # keep the fixture small: final stress and strain (+ final twin flags) only
        last = nincs - 1
        fields = {k: v for k, v in fields.items() if k in (f"sig_end_inc{last}", f"e_end_inc{last}", f"twinned_end_inc{last}")}
    if twin_hist:
        fields["twin_history"] = np.array(twin_hist)      # per increment: F_acc, F_eff, voxels reoriented
    c0v = S.C0m / np.outer(S.W, S.W)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
    np.savez_compressed(
        out, grid=np.array(grid), ngrains=ngrains, seed=seed, grain=ids, grain_rot=grot, c0_voigt=c0v,
        iudot=np.asarray(loading.iudot), udot=np.asarray(loading.udot), iscau=np.asarray(loading.iscau),
        scau=np.asarray(loading.scau), dt=dt, iters_per_inc=iters_per_inc, nincs=nincs,
        reports=np.array(reports), hcp=int(hcp), texture=int(texture), twinning=int(twinning), phase_kind=phase_kind, **fields)
    print("wrote", out, os.path.getsize(out), "bytes; last report", reports[-1][:5])


def main():
    lib = api.load_product()
    D = np.diag([-0.5, -0.5, 1.0])
    make_case(lib, "fcc8_strain", (8, 8, 8), 6, 0, api.Loading.strain_rate(D), 2e-4, 12, 2)
    make_case(lib, "fcc_12x10x8_tension", (12, 10, 8), 9, 3, api.Loading.uniaxial_tension(1.0), 2e-4, 10, 2)
    make_case(lib, "fcc_16x8x32_tension", (16, 8, 32), 12, 5, api.Loading.uniaxial_tension(1.0), 2e-4, 8, 2, slim=True)
    hcp = hcp_phase_np(with_twin=1, nrate=10.0, voce_mode=VOCE_HCP)
    make_case(lib, "hcp8_compression", (8, 8, 8), 5, 7, api.Loading.strain_rate(-D), 2e-4, 10, 2, phase=hcp, hcp=True)
    make_case(lib, "fcc8_texture", (8, 8, 8), 6, 2, api.Loading.strain_rate(D + np.array([[0, 0.3, 0], [-0.3, 0, 0], [0, 0, 0]])), 5e-4, 8, 3,
              texture=True)
    # HCP with easy tensile twinning and low PTR thresholds so that voxels reorient within a few increments
    # (same numbers as common_golden.twin_phase, which builds the product-side input from the product's own tables)
    twin = hcp_phase_np(with_twin=1, nrate=10.0, tau0_mode=(60.0, 120.0, 200.0, 25.0), voce_mode=VOCE_HCP, thr=(5.0e-8, 1.0e-13))
    make_case(lib, "hcp8_twin_texture", (8, 8, 8), 5, 11, api.Loading.strain_rate(D), 1e-3, 8, 4, phase=twin, hcp=True,
              texture=True, twinning=True, phase_kind=2)
    # PTR threshold that depends on F_eff / F_acc: with F_acc as the history sum voxels reorient in increments 2, 4, 6; with the
    # current-fraction sum (the defect fixed in round 2) they would in increments 2 and 5 only: the accumulated fraction must be the history sum
    twin2 = hcp_phase_np(with_twin=1, nrate=10.0, tau0_mode=(60.0, 120.0, 200.0, 25.0), voce_mode=VOCE_HCP, thr=(4.0e-8, 1.0e-12))
    make_case(lib, "hcp8_twin_ratio", (8, 8, 8), 5, 11, api.Loading.strain_rate(D), 1e-3, 6, 6, phase=twin2, hcp=True,
              texture=False, twinning=True, phase_kind=3, slim=True)


if __name__ == "__main__":
    main()
