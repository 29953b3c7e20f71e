"""CPU: the index arithmetic of the nz = 512 z kernel (k_zfused3, lapx_b200/csrc/kernels.cu: z3_forward / z3_inverse / z3_swap and
the Green stage's position -> frequency map), restated in numpy.  The kernel itself is tested on the GPU against the one-shot z
kernel and the oracle; this test pins WHY it is correct and conflict free:
  * in-place radix-8 passes at strides 64 / 8 / 1: decimation in frequency forward leaves frequency 64c + 8b + a at position 64a + 8b + c,
    decimation in time backward takes that order back to natural order;
  * rows z and z^1 are stored swapped where bit 3 of z is set between the first and the last pass; every thread still reads and
    writes its own eight elements (the first / last pass exchange with the neighbouring thread of the same warp only);
  * with the swap, every shared-memory access of every pass is bank-conflict free for the [z][4 kx] tile of 16-byte elements."""
import numpy as np

N = 512
W = np.exp(-2j * np.pi * np.arange(N) / N)
R8 = np.arange(8)


def dft8(v, inv=False):
    k = np.arange(8)
    return np.exp((2j if inv else -2j) * np.pi * np.outer(k, k) / 8) @ v


def swap(z):
    return z ^ ((z >> 3) & 1)


def forward(x):
    s = x.copy()
    t = s.copy()
    for u in range(64):                                   # stride 64, dense in, swapped out (after a __syncwarp)
        t[swap(u) + 64 * R8] = dft8(s[u + 64 * R8]) * W[(R8 * u) % N]
    s = t
    for u in range(64):                                   # stride 8 inside block u // 8
        idx = ((u >> 3) * 64 + (u & 7) + 8 * R8) ^ (R8 & 1)
        s[idx] = dft8(s[idx]) * W[(8 * (u & 7) * R8) % N]
    for u in range(64):                                   # stride 1
        idx = (8 * u + R8) ^ (u & 1)
        s[idx] = dft8(s[idx])
    return s


def inverse(s):
    t = s.copy()
    for u in range(64):
        idx = (8 * u + R8) ^ (u & 1)
        t[idx] = dft8(t[idx], True)
    for u in range(64):
        idx = ((u >> 3) * 64 + (u & 7) + 8 * R8) ^ (R8 & 1)
        t[idx] = dft8(t[idx] * np.conj(W[(8 * (u & 7) * R8) % N]), True)
    out = t.copy()
    for u in range(64):
        out[u + 64 * R8] = dft8(t[swap(u) + 64 * R8] * np.conj(W[(R8 * u) % N]), True)
    return out


def test_in_place_passes_give_the_digit_reversed_spectrum_and_invert():
    rng = np.random.default_rng(0)
    x = rng.normal(size=N) + 1j * rng.normal(size=N)
    s = forward(x)
    phys = np.arange(N)
    pos = swap(phys)                                      # Green stage: slot row -> position
    kz = ((pos & 7) << 6) | (pos & 0x38) | (pos >> 6)     # position -> frequency (base-8 digits reversed)
    assert np.abs(s - np.fft.fft(x)[kz]).max() < 1e-11
    assert np.abs(inverse(s) / N - x).max() < 1e-13
    assert sorted(kz) == list(range(N))


def test_every_thread_owns_its_elements_in_the_middle_passes():
    """Passes 2 and 3 (and their inverse counterparts) touch disjoint sets per thread: no barrier between loads and stores."""
    for sets in ([set((((u >> 3) * 64 + (u & 7) + 8 * R8) ^ (R8 & 1)).tolist()) for u in range(64)],
                 [set(((8 * u + R8) ^ (u & 1)).tolist()) for u in range(64)]):
        assert all(len(s) == 8 for s in sets)
        assert len(set().union(*sets)) == N
    # first pass: the swapped target rows of thread u are the source rows of thread u ^ bit3(u), a lane of the same warp (8 u per warp)
    for u in range(64):
        partner = swap(u)
        assert partner // 8 == u // 8


def test_all_passes_are_bank_conflict_free():
    """Lanes of a warp: t = u * 4 + col.  A quarter-warp (8 lanes) of 16-byte accesses must hit 8 distinct 16-byte bank groups."""
    def worst(rows):
        w = 1
        for warp in range(8):
            for qw in range(4):
                groups = []
                for lane in range(8 * qw, 8 * qw + 8):
                    u, col = 8 * warp + lane // 4, lane % 4
                    groups.append((rows(u) * 4 + col) & 7)
                w = max(w, max(groups.count(g) for g in set(groups)))
        return w
    for r in range(8):
        assert worst(lambda u: u + 64 * r) == 1                                          # dense rows (first load / last store)
        assert worst(lambda u: swap(u) + 64 * r) == 1                                    # swapped rows, stride 64
        assert worst(lambda u: ((u >> 3) * 64 + (u & 7) + 8 * r) ^ (r & 1)) == 1          # stride 8
        assert worst(lambda u: (8 * u + r) ^ (u & 1)) == 1                               # stride 1
        assert worst(lambda u: 8 * u + r) == 2                                           # ... which conflicts two-way without the swap
