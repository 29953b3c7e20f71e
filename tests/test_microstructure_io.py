"""CPU: per-voxel text microstructure reader / writer and the standalone tool.
Round trip of a Voronoi polycrystal through the text file: grain / phase ids bit exact, rotations to 1e-15 (the Bunge
angles are written with 17 significant digits)."""
import os
import subprocess

import numpy as np
import pytest

from lapx_b200 import build, microstructure as ms


def test_text_round_trip(product_lib, tmp_path):
    grid = (12, 10, 8)
    ids, grot = ms.voronoi(product_lib, grid, 17, 3)
    phase = (ids % 3 == 0).astype(np.int32)
    rot9 = ms.expand_rotations(ids, grot)
    path = tmp_path / "micro.txt"
    ms.write_txt(product_lib, path, ids, phase, rot9)
    lines = open(path).read().splitlines()
    assert len(lines) == ids.size and len(lines[0].split()) == 8
    g2, p2, r2 = ms.read_txt(product_lib, path, grid)
    assert np.array_equal(g2, ids) and np.array_equal(p2, phase)
    assert np.abs(r2 - rot9).max() < 1e-15
    # voxel lines in any order
    rng = np.random.default_rng(0)
    shuffled = tmp_path / "shuffled.txt"
    shuffled.write_text("\n".join(rng.permutation(lines)) + "\n")
    g3, p3, r3 = ms.read_txt(product_lib, shuffled, grid)
    assert np.array_equal(g3, ids) and np.array_equal(p3, phase) and np.array_equal(r3, r2)


def test_degenerate_euler_angles(product_lib, tmp_path):
    """Phi = 0 and Phi = 180 degrees (gimbal lock of the Bunge convention) survive the round trip."""
    ids = np.arange(4, dtype=np.int32).reshape(1, 2, 2)
    c, s = np.cos(0.3), np.sin(0.3)
    mats = np.array([np.eye(3), [[c, -s, 0], [s, c, 0], [0, 0, 1]], [[c, s, 0], [s, -c, 0], [0, 0, -1]], [[1, 0, 0], [0, -1, 0], [0, 0, -1]]], float)
    rot9 = ms.expand_rotations(ids, mats)
    ms.write_txt(product_lib, tmp_path / "d.txt", ids, None, rot9)
    g2, p2, r2 = ms.read_txt(product_lib, tmp_path / "d.txt", (2, 2, 1))
    assert np.array_equal(g2, ids) and not p2.any() and np.abs(r2 - rot9).max() < 1e-15


def test_reader_rejects_bad_files(product_lib, tmp_path):
    grid = (4, 4, 4)
    ids, grot = ms.voronoi(product_lib, grid, 3, 1)
    path = tmp_path / "m.txt"
    ms.write_txt(product_lib, path, ids, None, ms.expand_rotations(ids, grot))
    lines = open(path).read().splitlines()
    for name, body in [("short", lines[:-1]), ("dup", lines[:-1] + [lines[0]]), ("range", lines[:-1] + ["0 0 0 5 1 1 0 1"])]:
        f = tmp_path / (name + ".txt")
        f.write_text("\n".join(body) + "\n")
        with pytest.raises(OSError):
            ms.read_txt(product_lib, f, grid)
    with pytest.raises(OSError):
        ms.read_txt(product_lib, tmp_path / "absent.txt", grid)
    with pytest.raises(OSError):
        ms.read_txt(product_lib, path, (8, 4, 4))       # wrong grid


def test_standalone_tool(product_lib, tmp_path):
    build.build_driver()
    tool = os.path.join(os.path.dirname(build.OUT), "evpfft_microstructure")
    out = tmp_path / "v.txt"
    r = subprocess.run([tool, "voronoi", "8", "6", "4", "5", "9", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ids, grot = ms.voronoi(product_lib, (8, 6, 4), 5, 9)
    g2, p2, r2 = ms.read_txt(product_lib, out, (8, 6, 4))
    assert np.array_equal(g2, ids) and np.abs(r2 - ms.expand_rotations(ids, grot)).max() < 1e-15
    r = subprocess.run([tool, "info", str(out), "8", "6", "4"], capture_output=True, text=True)
    assert r.returncode == 0 and "192 voxels, 5 grains, 1 phases" in r.stdout
