"""Multi-GPU plumbing: one process per GPU, z-slab or pencil decomposition.

torch.distributed is used only as plumbing (rendezvous + broadcasting the NCCL unique id).  The FFT transposes are
done by the C library itself: by default (slabs) as TMA stores into the other ranks' buffers (CUDA IPC over NVLink)
fused into the y and z FFT kernels, or - evp_dist.transport = EVP_TRANSPORT_NCCL, EVP_TRANSPORT=nccl, pencils, or
when IPC mapping fails - as grouped ncclSend/ncclRecv all-to-alls on its communication stream.
The index helpers below restate the spectral-buffer layouts of lapx_b200/csrc/kernels.cuh
(SpecLayout) in numpy so that the decomposition logic can be tested on CPU with the gloo backend."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


def slab(nz: int, world: int, rank: int):
    """z-range [z0, z0+nzl) owned by `rank` (real space) — mirrored for ky in Fourier space."""
    if nz % world:
        raise ValueError("nz must be divisible by the number of ranks")
    nzl = nz // world
    return rank * nzl, nzl


class SpecLayout:
    """Half-spectrum work buffer [6 comps] x [nz] x [ny] x [nxp] complex, split over `world` ranks.

    y-split view (send side of the forward transpose, rows grouped by destination rank d = y // nyl):
        addr(c, zl, y)  = (y // nyl) * dstride + c * cstride + zl * zstride + (y % nyl) * nxp
    z-split view (receive side, planes grouped by source rank r = z // nzl):
        addr(c, z, yl)  = (z // nzl) * dstride + c * cstride + (z % nzl) * zstride + yl * nxp
    """

    def __init__(self, nx, ny, nz, world):
        self.nxh = nx // 2 + 1
        self.nxp = (self.nxh + 7) // 8 * 8
        self.nyl, self.nzl = ny // world, nz // world
        self.zstride = self.nyl * self.nxp
        self.cstride = self.nzl * self.zstride
        self.dstride = 6 * self.cstride
        self.world = world
        self.size = world * self.dstride          # complex elements per rank

    def row_ysplit(self, c, zl, y):
        return (y // self.nyl) * self.dstride + c * self.cstride + zl * self.zstride + (y % self.nyl) * self.nxp

    def row_zsplit(self, c, z, yl):
        return (z // self.nzl) * self.dstride + c * self.cstride + (z % self.nzl) * self.zstride + yl * self.nxp


class PencilLayout:
    """Pencil decomposition on a py x pz process grid, rank = iy*pz + iz (numpy restatement of evp_create's layouts).

    real space       rank (iy, iz) owns y in [iy*nyb, +nyb), z in [iz*nzl, +nzl), all x
    x stage (K2 out) kx pieces grouped by the row rank that will own them:
                         addr(c, row, k) = (k // kxl) * dx + c * cx + row * kxl + k % kxl,    row = zl*nyb + yb
    y stage          [nzl][ny][kxl] at kx0 = iy*kxl; x-side view: rows grouped by source row rank (nyb rows each),
                     z-side view: rows grouped by destination column rank (nyl = ny/pz rows each)
    z stage          [nz][nyl][kxl] at (ky0 = iz*nyl, kx0): planes grouped by source column rank
    Every stage holds 6 * nzl * ny * kxl complex elements per rank."""

    def __init__(self, nx, ny, nz, py, pz, rank):
        self.py, self.pz = py, pz
        self.iy, self.iz = rank // pz, rank % pz
        self.nxh = nx // 2 + 1
        self.kxl = ((self.nxh + py - 1) // py + 7) // 8 * 8
        self.kx0 = self.iy * self.kxl
        self.nxv = max(0, min(self.kxl, self.nxh - self.kx0))
        self.nyb, self.y0 = ny // py, self.iy * (ny // py)
        self.nzl, self.z0 = nz // pz, self.iz * (nz // pz)
        self.nyl, self.ky0 = ny // pz, self.iz * (ny // pz)
        self.cx = self.nzl * self.nyb * self.kxl          # x-side: component stride, group stride = 6*cx
        self.dx = 6 * self.cx
        self.cz = self.nzl * self.nyl * self.kxl          # z-side
        self.dz = 6 * self.cz
        self.size = 6 * self.nzl * ny * self.kxl

    def addr_x(self, c, row, k):
        return (k // self.kxl) * self.dx + c * self.cx + row * self.kxl + k % self.kxl

    def row_xside(self, c, zl, y):
        """y stage, x-side view (what the row exchange delivers / what K5 writes)."""
        return (y // self.nyb) * self.dx + c * self.cx + zl * self.nyb * self.kxl + (y % self.nyb) * self.kxl

    def row_zside(self, c, zl, y):
        """y stage, z-side view (what K3 writes / the column exchange back delivers)."""
        return (y // self.nyl) * self.dz + c * self.cz + zl * self.nyl * self.kxl + (y % self.nyl) * self.kxl

    def row_zstage(self, c, z, yl):
        return (z // self.nzl) * self.dz + c * self.cz + (z % self.nzl) * self.nyl * self.kxl + yl * self.kxl

    def row_group(self):
        """Ranks of my row (same iz, iy = 0..py-1) in exchange order."""
        return [j * self.pz + self.iz for j in range(self.py)]

    def col_group(self):
        return [self.iy * self.pz + j for j in range(self.pz)]


def nccl_unique_id(lib, rank: int, broadcast_bytes) -> "C.Array":
    """Rank 0 asks the library for an ncclUniqueId; `broadcast_bytes(np.uint8[128]) -> np.uint8[128]` ships it."""
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        rc = lib.evp_nccl_unique_id(buf)
        if rc != 0:
            raise api.EvpError(rc, "evp_nccl_unique_id failed")
    arr = broadcast_bytes(np.frombuffer(buf, dtype=np.uint8).copy())
    return (C.c_uint8 * 128)(*[int(v) for v in arr])


def make_dist(lib, world: int, rank: int, device: int, td=None, transport: int = api.TRANSPORT_AUTO, py: int = 1) -> api.Dist | None:
    """Build the evp_dist argument; `td` is an initialised torch.distributed module (any backend).
    py >= 2 selects the pencil decomposition on a py x (world/py) process grid."""
    if world == 1:
        return None
    import torch

    def bcast(a):
        use_cuda = td.get_backend() == "nccl"
        t = torch.from_numpy(a.copy())
        if use_cuda:
            t = t.cuda()
        td.broadcast(t, 0)
        return t.cpu().numpy()

    uid = nccl_unique_id(lib, rank, bcast)
    return api.Dist(world, rank, device, int(transport), uid, int(py))
