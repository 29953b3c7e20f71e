"""Build the product library lapx_b200/libevpfft_b200.so (sm_100a only) and the CPU oracle.

    python -m lapx_b200.build            # both
The .so files are git-ignored but travel to the GPU box with the gpurun snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libevpfft_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xptxas=-v",
    "-Xcompiler", "-fPIC,-O3,-fopenmp,-Wall", "-shared", "-cudart", "shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    srcs = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cpp")):
            srcs.append(os.path.join(CSRC, f))
    return srcs


def needs_build(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build_product(force=False, verbose=False):
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    deps.append(os.path.join(ROOT, "include", "evpfft.h"))
    if not force and not needs_build(OUT, deps):
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + ["-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
                                    "-I", os.path.join(ROOT, "include"), "-o", OUT] + sources() + ["-lgomp", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "build_ptxas.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed, see " + log)
    return OUT


def build_driver(force=False):
    """Host driver executable (C++17) linked against the product library."""
    src = os.path.join(CSRC, "driver", "evpfft_main.cpp")
    out = os.path.join(HERE, "evpfft_driver")
    if force or needs_build(out, [src, OUT, os.path.join(ROOT, "include", "evpfft.h")]):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-o", out, src, "-L", HERE, "-levpfft_b200", "-Wl,-rpath,$ORIGIN"], check=True)
    return out


def build_oracle(force=False):
    out = os.path.join(ROOT, "oracle", "libevp_oracle.so")
    deps = [os.path.join(ROOT, "oracle", "evp_oracle.cpp"), os.path.join(ROOT, "include", "evpfft.h")]
    if force or needs_build(out, deps):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return out


if __name__ == "__main__":
    print(build_product(force="--force" in sys.argv, verbose=True))
    print(build_driver(force="--force" in sys.argv))
    print(build_oracle(force="--force" in sys.argv))
