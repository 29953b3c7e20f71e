"""CPU, world_size 2 over gloo: the N>1 host-side logic — slab ownership, per-rank microstructure
generation, NCCL-id style broadcast plumbing, and the send/recv spectral layouts of the FFT
transpose (numpy restatement of SpecLayout) reproducing a global rfftn through a real all-to-all."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lapx_b200 import api, distributed as dist, microstructure as ms
        lib = api.load_product()
        nx, ny, nz = 16, 8, 12
        # 1. slab ownership + per-rank Voronoi == slice of the global tessellation
        z0, nzl = dist.slab(nz, world, rank)
        ids, grot = ms.voronoi(lib, (nx, ny, nz), 20, 3, z0=z0, nzl=nzl)
        full, grot_full = ms.voronoi(lib, (nx, ny, nz), 20, 3)
        assert np.array_equal(ids, full[z0:z0 + nzl]) and np.array_equal(grot, grot_full)
        gathered = [torch.zeros(ids.shape, dtype=torch.int32) for _ in range(world)]
        td.all_gather(gathered, torch.from_numpy(ids))
        assert np.array_equal(np.concatenate([g.numpy() for g in gathered]), full)
        # 2. id broadcast plumbing (the payload stands in for the ncclUniqueId)
        payload = np.arange(128, dtype=np.uint8) if rank == 0 else np.zeros(128, np.uint8)
        t = torch.from_numpy(payload.copy())
        td.broadcast(t, 0)
        assert np.array_equal(t.numpy(), np.arange(128, dtype=np.uint8))
        # 3. transposed FFT through the send / recv layouts and a real all-to-all
        rng = np.random.default_rng(7)
        field = rng.normal(size=(6, nz, ny, nx))          # same on both ranks (same seed)
        L = dist.SpecLayout(nx, ny, nz, world)
        local = field[:, z0:z0 + nzl]
        xy = np.fft.fft(np.fft.rfft(local, axis=3), axis=2)          # K2 + K3 on the local slab
        send = np.zeros(L.size, complex)
        for c in range(6):
            for zl in range(nzl):
                for y in range(ny):
                    o = L.row_ysplit(c, zl, y)
                    send[o:o + L.nxh] = xy[c, zl, y]
        recv = np.zeros(L.size, complex)
        st = torch.from_numpy(np.ascontiguousarray(send.view(np.float64)))
        rt = torch.from_numpy(recv.view(np.float64))
        td.all_to_all_single(rt, st)                                  # equal contiguous chunks, as ncclSend/Recv
        recv = rt.numpy().view(np.complex128)
        ky0 = rank * L.nyl
        ref = np.fft.fftn(np.fft.rfft(field, axis=3), axes=(1, 2))    # global spectrum
        for c in range(6):
            for yl in range(L.nyl):
                col = np.stack([recv[L.row_zsplit(c, z, yl):L.row_zsplit(c, z, yl) + L.nxh] for z in range(nz)])
                got = np.fft.fft(col, axis=0)                         # K4's z transform
                assert np.abs(got - ref[c, :, ky0 + yl, :]).max() < 1e-11
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        td.destroy_process_group()


def test_slab_decomposition_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
