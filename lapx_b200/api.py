"""ctypes binding of include/evpfft.h — the host-side mirror of the C ABI.

The reference's own interface for this path is absent (/root/reference holds only LICENSE,
see SURVEY.md §0); the names, argument meaning and error behaviour here follow the C ABI
proposed in SURVEY.md §8(b) one to one, so that parity tests drive the CUDA library and the
CPU oracle through the very same calls.

`load_product()` loads lapx_b200/libevpfft_b200.so and raises if it is missing: there is no
CPU fallback in the product path.  The oracle is loaded only by tests / smoke / bench baselines
through `load_library(path)`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

EVP_MAX_SYS = 32
EVP_MAX_MODES = 8
EVP_MAX_PHASES = 4

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "libevpfft_b200.so")

# evp_field
FIELD_STRESS, FIELD_STRAIN, FIELD_PLASTIC_STRAIN, FIELD_PLASTIC_RATE = 0, 1, 2, 3
FIELD_CRSS, FIELD_ROTATION, FIELD_GRAIN, FIELD_PHASE, FIELD_GAMMA_ACC = 4, 5, 6, 7, 8
FIELD_TWIN_FRACTION, FIELD_STRAIN_INCR, FIELD_LOCAL_ROTATION, FIELD_TWINNED = 9, 10, 11, 12
_INT_FIELDS = (FIELD_GRAIN, FIELD_PHASE, FIELD_TWINNED)
# evp_transport_kind
TRANSPORT_AUTO, TRANSPORT_NCCL, TRANSPORT_P2P = 0, 1, 2


class EvpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"evp error {code}: {msg}")
        self.code = code


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double)]


class Phase(C.Structure):
    _fields_ = [
        ("nsys", C.c_int32), ("nmodes", C.c_int32),
        ("c_voigt", C.c_double * 36),
        ("b", (C.c_double * 3) * EVP_MAX_SYS),
        ("n", (C.c_double * 3) * EVP_MAX_SYS),
        ("mode", C.c_int32 * EVP_MAX_SYS),
        ("twin", C.c_int32 * EVP_MAX_MODES),
        ("gamma0", C.c_double * EVP_MAX_MODES),
        ("nrate", C.c_double * EVP_MAX_MODES),
        ("tau0", C.c_double * EVP_MAX_MODES),
        ("tau1", C.c_double * EVP_MAX_MODES),
        ("theta0", C.c_double * EVP_MAX_MODES),
        ("theta1", C.c_double * EVP_MAX_MODES),
        ("hlat", (C.c_double * EVP_MAX_MODES) * EVP_MAX_MODES),
        ("twin_shear", C.c_double * EVP_MAX_MODES),
        ("twin_thr1", C.c_double), ("twin_thr2", C.c_double),
    ]


class Dist(C.Structure):
    """evp_dist: py <= 1 = z-slabs over nranks; py >= 2 = pencils on a py x (nranks/py) process grid."""
    _fields_ = [("nranks", C.c_int32), ("rank", C.c_int32), ("device", C.c_int32),
                ("transport", C.c_int32), ("nccl_id", C.c_uint8 * 128),
                ("py", C.c_int32), ("reserved", C.c_int32 * 3)]


class Ctrl(C.Structure):
    _fields_ = [("tol_stress", C.c_double), ("tol_strain", C.c_double),
                ("itmax", C.c_int32), ("itmin", C.c_int32),
                ("tol_newton", C.c_double), ("newton_itmax", C.c_int32),
                ("update_texture", C.c_int32), ("update_twinning", C.c_int32)]


class IterReport(C.Structure):
    _fields_ = [("iter", C.c_int32), ("newton_max", C.c_int32), ("newton_mean", C.c_double),
                ("err_stress", C.c_double), ("err_strain", C.c_double),
                ("savg", C.c_double * 6), ("emacro", C.c_double * 6),
                ("converged", C.c_int32), ("nonfinite", C.c_int32), ("unconverged", C.c_int64)]


class StepReport(C.Structure):
    _fields_ = [("iters", C.c_int32), ("converged", C.c_int32),
                ("err_stress", C.c_double), ("err_strain", C.c_double),
                ("savg", C.c_double * 6), ("emacro", C.c_double * 6), ("epavg", C.c_double * 6),
                ("seconds", C.c_double), ("twin_acc", C.c_double), ("twin_eff", C.c_double), ("reoriented", C.c_int64)]


# every symbol include/evpfft.h declares; tests check both libraries against this list
ABI_SYMBOLS_COMMON = [
    "evp_abi_version", "evp_backend", "evp_create", "evp_destroy", "evp_last_error",
    "evp_local_slab", "evp_nsys_max", "evp_set_microstructure", "evp_set_reference_medium",
    "evp_get_reference_medium", "evp_set_control", "evp_set_loading", "evp_begin_increment",
    "evp_equilibrium_iter", "evp_op_green", "evp_op_constitutive", "evp_end_increment",
    "evp_step", "evp_equilibrium_iters", "evp_field_components", "evp_get_field",
    "evp_set_field", "evp_get_macro", "evp_debug_spectrum", "evp_stream",
    "evp_set_profiling", "evp_last_kernel_ms", "evp_transport", "evp_save_state", "evp_load_state",
    "evp_local_block", "evp_build_id", "evp_launch_count", "evp_debug_fp64_peak",
]
ABI_SYMBOLS_PRODUCT_ONLY = ["evp_phase_fcc", "evp_phase_hcp", "evp_voronoi", "evp_nccl_unique_id",
                            "evp_read_microstructure_txt", "evp_write_microstructure_txt"]


def _proto(lib):
    """Attach argtypes/restype.  Missing symbols are skipped here (the ABI-completeness test checks
    them explicitly); calling a missing one raises AttributeError, never falls back."""
    H = C.c_void_p
    P = C.POINTER
    protos = {
        "evp_abi_version": ([], C.c_int),
        "evp_backend": ([], C.c_char_p),
        "evp_create": ([P(Grid), P(Phase), C.c_int32, P(Dist), P(H)], C.c_int),
        "evp_destroy": ([H], C.c_int),
        "evp_last_error": ([H], C.c_char_p),
        "evp_local_slab": ([H, P(C.c_int32), P(C.c_int32)], C.c_int),
        "evp_nsys_max": ([H], C.c_int),
        "evp_set_microstructure": ([H, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "evp_set_reference_medium": ([H, C.c_void_p], C.c_int),
        "evp_get_reference_medium": ([H, C.c_void_p], C.c_int),
        "evp_set_control": ([H, P(Ctrl)], C.c_int),
        "evp_set_loading": ([H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "evp_begin_increment": ([H, C.c_double], C.c_int),
        "evp_equilibrium_iter": ([H, P(IterReport)], C.c_int),
        "evp_op_green": ([H], C.c_int),
        "evp_op_constitutive": ([H, P(IterReport)], C.c_int),
        "evp_end_increment": ([H, P(StepReport)], C.c_int),
        "evp_step": ([H, C.c_double, P(StepReport)], C.c_int),
        "evp_equilibrium_iters": ([H, C.c_int32, P(IterReport)], C.c_int),
        "evp_field_components": ([H, C.c_int], C.c_int),
        "evp_get_field": ([H, C.c_int, C.c_void_p, C.c_size_t], C.c_int),
        "evp_set_field": ([H, C.c_int, C.c_void_p, C.c_size_t], C.c_int),
        "evp_get_macro": ([H, C.c_void_p, C.c_void_p], C.c_int),
        "evp_debug_spectrum": ([H, C.c_int32, C.c_void_p], C.c_int),
        "evp_stream": ([H], C.c_void_p),
        "evp_set_profiling": ([H, C.c_int32], C.c_int),
        "evp_last_kernel_ms": ([H, C.c_void_p], C.c_int),
        "evp_transport": ([H], C.c_int),
        "evp_save_state": ([H, C.c_char_p], C.c_int),
        "evp_load_state": ([H, C.c_char_p], C.c_int),
        "evp_phase_fcc": ([P(Phase)] + [C.c_double] * 9, C.c_int),
        "evp_phase_hcp": ([P(Phase), C.c_double, C.c_void_p, C.c_int32, C.c_double, C.c_double,
                           C.c_void_p, C.c_void_p], C.c_int),
        "evp_voronoi": ([P(Grid), C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p], C.c_int),
        "evp_nccl_unique_id": ([C.c_void_p], C.c_int),
        "evp_local_block": ([H, P(C.c_int32), P(C.c_int32), P(C.c_int32), P(C.c_int32)], C.c_int),
        "evp_build_id": ([], C.c_char_p),
        "evp_launch_count": ([H], C.c_int64),
        "evp_debug_fp64_peak": ([H, C.c_int32, P(C.c_double)], C.c_int),
        "evp_read_microstructure_txt": ([C.c_char_p, P(Grid), C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "evp_write_microstructure_txt": ([C.c_char_p, P(Grid), C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "evp_oracle_threads": ([], C.c_int),
        "evp_oracle_set_threads": ([C.c_int], C.c_int),
    }
    for name, (args, res) in protos.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            continue
        fn.argtypes = args
        fn.restype = res
    return lib


_LIBS = {}


def load_library(path: str):
    """Load one implementation of the ABI (RTLD_LOCAL, so both can coexist in a process)."""
    path = os.path.abspath(path)
    if path not in _LIBS:
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing. Build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is deliberately no CPU fallback for the product path.")
        _LIBS[path] = _proto(C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | os.RTLD_NOW))
    return _LIBS[path]


def load_product():
    return load_library(PRODUCT_LIB)


def _arr(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, a.ctypes.data_as(C.c_void_p)


@dataclass
class Loading:
    """Mixed boundary conditions, see evp_set_loading."""
    iudot: np.ndarray
    udot: np.ndarray
    iscau: np.ndarray
    scau: np.ndarray

    @staticmethod
    def uniaxial_tension(rate: float, axis: int = 2) -> "Loading":
        """Imposed L_aa = rate, shear velocity gradients 0, lateral normal stresses 0."""
        iudot = np.ones((3, 3), np.int32)
        udot = np.zeros((3, 3))
        iscau = np.zeros(6, np.int32)
        for k in range(3):
            if k != axis:
                iudot[k, k] = 0
                iscau[k] = 1
        udot[axis, axis] = rate
        return Loading(iudot, udot, iscau, np.zeros(6))

    @staticmethod
    def strain_rate(D: np.ndarray) -> "Loading":
        """Fully imposed velocity gradient D (3x3)."""
        return Loading(np.ones((3, 3), np.int32), np.asarray(D, float).reshape(3, 3),
                       np.zeros(6, np.int32), np.zeros(6))

    @staticmethod
    def plane_strain_compression(rate: float) -> "Loading":
        """L_33 = -rate, L_22 = 0 (constrained), sigma_11 = 0 (free extension along 1)."""
        iudot = np.ones((3, 3), np.int32)
        udot = np.zeros((3, 3))
        iscau = np.zeros(6, np.int32)
        iudot[0, 0] = 0
        iscau[0] = 1
        udot[2, 2] = -rate
        return Loading(iudot, udot, iscau, np.zeros(6))


class Solver:
    """One handle of the C ABI.  Method names drop the evp_ prefix and nothing else."""

    def __init__(self, lib, grid: Sequence[int], phases: Sequence[Phase], spacing=(1.0, 1.0, 1.0),
                 dist: Optional[Dist] = None):
        self.lib = lib
        nx, ny, nz = (int(v) for v in grid)
        self.grid = Grid(nx, ny, nz, *[float(s) for s in spacing])
        arr = (Phase * len(phases))(*phases)
        h = C.c_void_p()
        rc = lib.evp_create(C.byref(self.grid), arr, len(phases), C.byref(dist) if dist is not None else None,
                            C.byref(h))
        if rc != 0:
            raise EvpError(rc, (lib.evp_last_error(None) or b"").decode())
        self.h = h
        y0, nyl, z0, nzl = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._check(lib.evp_local_block(h, C.byref(y0), C.byref(nyl), C.byref(z0), C.byref(nzl)))
        self.y0, self.nyl, self.z0, self.nzl = y0.value, nyl.value, z0.value, nzl.value   # local block (slab: nyl = ny)
        self.nx, self.ny, self.nz = nx, ny, nz
        self.nlocal = nx * self.nyl * self.nzl
        self.nsys_max = lib.evp_nsys_max(h)

    # -- plumbing ------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise EvpError(rc, (self.lib.evp_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.evp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def backend(self) -> str:
        return self.lib.evp_backend().decode()

    # -- set-up --------------------------------------------------------------------------
    def set_microstructure(self, grain, phase, rot9):
        n = self.nlocal
        g, gp = _arr(grain, np.int32)
        r, rp = _arr(rot9, np.float64)
        if g.size != n or r.size != 9 * n:
            raise EvpError(-1, "set_microstructure: array size does not match the local slab")
        if phase is None:
            pp = None
        else:
            p, pp = _arr(phase, np.int32)
            if p.size != n:
                raise EvpError(-1, "set_microstructure: phase size mismatch")
        self._check(self.lib.evp_set_microstructure(self.h, gp, pp, rp))

    def set_reference_medium(self, c0_voigt=None):
        if c0_voigt is None:
            self._check(self.lib.evp_set_reference_medium(self.h, None))
        else:
            c, cp = _arr(c0_voigt, np.float64)
            assert c.size == 36
            self._check(self.lib.evp_set_reference_medium(self.h, cp))

    def get_reference_medium(self) -> np.ndarray:
        out = np.zeros(36)
        self._check(self.lib.evp_get_reference_medium(self.h, out.ctypes.data_as(C.c_void_p)))
        return out.reshape(6, 6)

    def set_control(self, tol_stress=1e-6, tol_strain=1e-6, itmax=100, itmin=1, tol_newton=1e-6,
                    newton_itmax=100, update_texture=0, update_twinning=0):
        c = Ctrl(tol_stress, tol_strain, itmax, itmin, tol_newton, newton_itmax, int(update_texture), int(update_twinning))
        self._check(self.lib.evp_set_control(self.h, C.byref(c)))

    def set_loading(self, ld: Loading):
        a, ap = _arr(ld.iudot, np.int32)
        b, bp = _arr(ld.udot, np.float64)
        c, cp = _arr(ld.iscau, np.int32)
        d, dp = _arr(ld.scau, np.float64)
        self._check(self.lib.evp_set_loading(self.h, ap, bp, cp, dp))

    # -- hot path ------------------------------------------------------------------------
    def begin_increment(self, dt: float):
        self._check(self.lib.evp_begin_increment(self.h, float(dt)))

    def equilibrium_iter(self) -> IterReport:
        r = IterReport()
        self._check(self.lib.evp_equilibrium_iter(self.h, C.byref(r)))
        return r

    def equilibrium_iters(self, n: int) -> IterReport:
        r = IterReport()
        self._check(self.lib.evp_equilibrium_iters(self.h, int(n), C.byref(r)))
        return r

    def op_green(self):
        self._check(self.lib.evp_op_green(self.h))

    def op_constitutive(self) -> IterReport:
        r = IterReport()
        self._check(self.lib.evp_op_constitutive(self.h, C.byref(r)))
        return r

    def end_increment(self) -> StepReport:
        r = StepReport()
        self._check(self.lib.evp_end_increment(self.h, C.byref(r)))
        return r

    def step(self, dt: float) -> StepReport:
        r = StepReport()
        self._check(self.lib.evp_step(self.h, float(dt), C.byref(r)))
        return r

    # -- fields --------------------------------------------------------------------------
    def field_components(self, f: int) -> int:
        return self.lib.evp_field_components(self.h, f)

    def get_field(self, f: int) -> np.ndarray:
        nc = self.field_components(f)
        dt = np.int32 if f in _INT_FIELDS else np.float64
        out = np.empty((nc, self.nzl, self.nyl, self.nx), dtype=dt)
        self._check(self.lib.evp_get_field(self.h, f, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def set_field(self, f: int, a):
        nc = self.field_components(f)
        dt = np.int32 if f in _INT_FIELDS else np.float64
        a = np.ascontiguousarray(a, dtype=dt)
        if a.size != nc * self.nlocal:
            raise EvpError(-1, "set_field: size mismatch")
        self._check(self.lib.evp_set_field(self.h, f, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def get_macro(self):
        e, s = np.zeros(6), np.zeros(6)
        self._check(self.lib.evp_get_macro(self.h, e.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p)))
        return e, s

    def debug_spectrum(self, comp: int) -> np.ndarray:
        out = np.empty((self.nz, self.ny, self.nx // 2 + 1), dtype=np.complex128)
        self._check(self.lib.evp_debug_spectrum(self.h, comp, out.ctypes.data_as(C.c_void_p)))
        return out

    def save_state(self, path: str):
        self._check(self.lib.evp_save_state(self.h, str(path).encode()))

    def load_state(self, path: str):
        self._check(self.lib.evp_load_state(self.h, str(path).encode()))

    def set_profiling(self, on: bool):
        self._check(self.lib.evp_set_profiling(self.h, int(on)))

    def last_kernel_ms(self) -> np.ndarray:
        ms = np.zeros(8)
        self._check(self.lib.evp_last_kernel_ms(self.h, ms.ctypes.data_as(C.c_void_p)))
        return ms

    def stream(self) -> int:
        return int(self.lib.evp_stream(self.h) or 0)

    def launch_count(self) -> int:
        return int(self.lib.evp_launch_count(self.h))

    def fp64_peak_tflops(self, reps: int = 5) -> float:
        v = C.c_double()
        self._check(self.lib.evp_debug_fp64_peak(self.h, int(reps), C.byref(v)))
        return float(v.value)

    def transport(self) -> str:
        return "p2p" if self.lib.evp_transport(self.h) == TRANSPORT_P2P else "nccl"
