"""ORACLE (test infrastructure only): brute-force numpy restatement of the integer Voronoi rule of
SURVEY.md §8(d) — voxel centre 2i+1 and seeds on the 2x refined lattice, squared periodic distance
in int64, ties to the lowest grain id, seeds from splitmix64(seed).  Used to check grain indexing
bit-exactly; the reference has no counterpart (mount holds only LICENSE)."""
import numpy as np

_M = (1 << 64) - 1


def splitmix64_stream(seed: int, n: int):
    out, st = [], seed & _M
    for _ in range(n):
        st = (st + 0x9E3779B97F4A7C15) & _M
        z = st
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
        out.append(z ^ (z >> 31))
    return out


def voronoi_ids(nx: int, ny: int, nz: int, ngrains: int, seed: int = 0) -> np.ndarray:
    draws = splitmix64_stream(seed, 3 * ngrains)
    L = np.array([2 * nx, 2 * ny, 2 * nz], dtype=np.int64)
    s = np.array([[draws[3 * i + k] % int(L[k]) for k in range(3)] for i in range(ngrains)], dtype=np.int64)
    px = 2 * np.arange(nx, dtype=np.int64) + 1
    py = 2 * np.arange(ny, dtype=np.int64) + 1
    pz = 2 * np.arange(nz, dtype=np.int64) + 1
    best = np.full((nz, ny, nx), np.iinfo(np.int64).max, dtype=np.int64)
    ids = np.full((nz, ny, nx), -1, dtype=np.int32)
    for g in range(ngrains):
        dx = np.abs(px - s[g, 0]); dx = np.minimum(dx, L[0] - dx)
        dy = np.abs(py - s[g, 1]); dy = np.minimum(dy, L[1] - dy)
        dz = np.abs(pz - s[g, 2]); dz = np.minimum(dz, L[2] - dz)
        d2 = dz[:, None, None] ** 2 + dy[None, :, None] ** 2 + dx[None, None, :] ** 2
        m = d2 < best                      # strict: ties keep the lower id
        best[m] = d2[m]
        ids[m] = g
    return ids
