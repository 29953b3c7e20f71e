"""CPU, gloo, world 2 (2x1) and 4 (2x2): the host-side logic of the PENCIL decomposition — block ownership, the x-stage /
y-stage / z-stage layouts of evp_create (numpy restatement: lapx_b200.distributed.PencilLayout) and the row / column
all-to-alls — reproduces the global rfftn and takes it back to real space."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _a2a(buf, piece, group_ranks, group):
    import torch
    send = torch.from_numpy(np.ascontiguousarray(buf.view(np.float64)))
    recv = torch.zeros_like(send)
    td.all_to_all_single(recv, send, group=group)        # equal contiguous pieces, as the grouped ncclSend/ncclRecv
    return recv.numpy().view(np.complex128)


def _worker(rank, world, py, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lapx_b200 import api, distributed as dist, microstructure as ms
        lib = api.load_product()
        pz = world // py
        nx, ny, nz = 20, 8, 12 if pz != 4 else 16
        L = dist.PencilLayout(nx, ny, nz, py, pz, rank)
        # every rank must create every group (torch.distributed rule), in the same order
        rows = {iz: td.new_group([j * pz + iz for j in range(py)]) for iz in range(pz)}
        cols = {iy: td.new_group([iy * pz + j for j in range(pz)]) for iy in range(py)}
        grow, gcol = rows[L.iz], cols[L.iy]
        # 1. block ownership: per-rank Voronoi block == slice of the global tessellation
        full, _ = ms.voronoi(lib, (nx, ny, nz), 20, 3)

        class S:  # what ms.voronoi_block needs from a Solver
            y0, nyl, z0, nzl = L.y0, L.nyb, L.z0, L.nzl
        ids, _ = ms.voronoi_block(lib, (nx, ny, nz), 20, 3, S)
        assert np.array_equal(ids, full[L.z0:L.z0 + L.nzl, L.y0:L.y0 + L.nyb, :])
        # 2. forward transform through the three layouts
        rng = np.random.default_rng(7)
        field = rng.normal(size=(6, nz, ny, nx))          # same on all ranks
        local = field[:, L.z0:L.z0 + L.nzl, L.y0:L.y0 + L.nyb, :]
        fx = np.fft.rfft(local, axis=3)                   # K2
        A = np.zeros(L.size, complex)
        for c in range(6):
            for zl in range(L.nzl):
                for yb in range(L.nyb):
                    for k in range(L.nxh):
                        A[L.addr_x(c, zl * L.nyb + yb, k)] = fx[c, zl, yb, k]
        B = _a2a(A, L.dx, L.row_group(), grow)            # row exchange: piece j -> rank (j, iz)
        ystage = np.zeros((6, L.nzl, ny, L.kxl), complex)
        for c in range(6):
            for zl in range(L.nzl):
                for y in range(ny):
                    o = L.row_xside(c, zl, y)
                    ystage[c, zl, y] = B[o:o + L.kxl]
        fy = np.fft.fft(ystage, axis=2)                   # K3
        A = np.zeros(L.size, complex)
        for c in range(6):
            for zl in range(L.nzl):
                for y in range(ny):
                    o = L.row_zside(c, zl, y)
                    A[o:o + L.kxl] = fy[c, zl, y]
        B = _a2a(A, L.dz, L.col_group(), gcol)            # column exchange: piece j -> rank (iy, j)
        zstage = np.zeros((6, nz, L.nyl, L.kxl), complex)
        for c in range(6):
            for z in range(nz):
                for yl in range(L.nyl):
                    o = L.row_zstage(c, z, yl)
                    zstage[c, z, yl] = B[o:o + L.kxl]
        fz = np.fft.fft(zstage, axis=1)                   # K4 forward
        ref = np.fft.fftn(np.fft.rfft(field, axis=3), axes=(1, 2))
        assert np.abs(fz[..., :L.nxv] - ref[:, :, L.ky0:L.ky0 + L.nyl, L.kx0:L.kx0 + L.nxv]).max() < 1e-11
        # 3. and back: inverse z, column exchange, inverse y, row exchange, inverse x
        bz = np.fft.ifft(fz, axis=1)
        A = np.zeros(L.size, complex)
        for c in range(6):
            for z in range(nz):
                for yl in range(L.nyl):
                    o = L.row_zstage(c, z, yl)
                    A[o:o + L.kxl] = bz[c, z, yl]
        B = _a2a(A, L.dz, L.col_group(), gcol)
        for c in range(6):
            for zl in range(L.nzl):
                for y in range(ny):
                    o = L.row_zside(c, zl, y)
                    ystage[c, zl, y] = B[o:o + L.kxl]
        by = np.fft.ifft(ystage, axis=2)
        A = np.zeros(L.size, complex)
        for c in range(6):
            for zl in range(L.nzl):
                for y in range(ny):
                    o = L.row_xside(c, zl, y)
                    A[o:o + L.kxl] = by[c, zl, y]
        B = _a2a(A, L.dx, L.row_group(), grow)
        xs = np.zeros((6, L.nzl, L.nyb, L.nxh), complex)
        for c in range(6):
            for zl in range(L.nzl):
                for yb in range(L.nyb):
                    for k in range(L.nxh):
                        xs[c, zl, yb, k] = B[L.addr_x(c, zl * L.nyb + yb, k)]
        back = np.fft.irfft(xs, n=nx, axis=3)
        assert np.abs(back - local).max() < 1e-12
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()))
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world,py", [(2, 2), (4, 2)])
def test_pencil_decomposition_gloo(world, py):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, py, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
