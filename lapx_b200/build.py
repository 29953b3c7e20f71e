"""Build the product library lapx_b200/libevpfft_b200.so (sm_100a only), the host driver and the CPU oracle.

    python -m lapx_b200.build [--force]

Staleness is decided by CONTENT, not by mtime: the sha256 of every source the library is compiled from (plus the
compiler flags) is embedded in the binary (`evp_build_id()`, marker "EVPSRC:<hex>") and compared with the digest of the
sources in the tree.  A binary that travelled with a gpurun snapshot is therefore either provably the build of the
sources next to it, or it is rebuilt.  The oracle is compiled -march=native, so its marker also carries a signature of
the host CPU: on another machine it is rebuilt before use.
The .so files are git-ignored but travel to the GPU box with the gpurun snapshot."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libevpfft_b200.so")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_OUT = os.path.join(ORACLE_DIR, "libevp_oracle.so")

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


def _cxx() -> str:
    # the image exports CXX=/opt/gcc/bin/g++, a wrapper without libgomp.spec; use the distro compiler
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cpp"))]


def product_deps():
    deps = sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h", ".hpp"))]
    deps.append(os.path.join(ROOT, "include", "evpfft.h"))
    return deps


def digest(paths, extra=()) -> str:
    h = hashlib.sha256()
    for p in paths:
        h.update(os.path.basename(p).encode() + b"\0")
        with open(p, "rb") as f:
            h.update(f.read())
        h.update(b"\0")
    for e in extra:
        h.update(str(e).encode() + b"\0")
    return h.hexdigest()[:24]


def source_id() -> str:
    """Digest of the sources + flags the product library is (to be) compiled from."""
    return digest(product_deps(), NVCC_FLAGS)


def has_marker(path: str, marker: str) -> bool:
    if not os.path.exists(path):
        return False
    with open(path, "rb") as f:
        return marker.encode() in f.read()


def needs_build(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build_product(force=False, verbose=False):
    sid = "EVPSRC:" + source_id()
    if not force and has_marker(OUT, sid):
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + ["-ccbin", _cxx(), f'-DEVP_SRC_HASH="{sid}"',
                                    "-I", os.path.join(ROOT, "include"), "-o", OUT] + sources() + ["-lgomp", "-ldl", "-lpthread"]
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
    """Host driver executable and the microstructure tool (C++17) linked against the product library."""
    outs = []
    for name, srcname in (("evpfft_driver", "evpfft_main.cpp"), ("evpfft_microstructure", "microstructure_tool.cpp")):
        src = os.path.join(CSRC, "driver", srcname)
        if not os.path.exists(src):
            continue
        out = os.path.join(HERE, name)
        if force or needs_build(out, [src, OUT, os.path.join(ROOT, "include", "evpfft.h")]):
            subprocess.run([_cxx(), "-O2", "-std=c++17", "-o", out, src, "-L", HERE, "-levpfft_b200", "-Wl,-rpath,$ORIGIN"], check=True)
        outs.append(out)
    return outs[0]


def host_signature() -> str:
    """What -march=native depends on: the CPU model and its feature flags."""
    model, flags = "", ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name") and not model:
                    model = line.split(":", 1)[1].strip()
                elif line.startswith("flags") and not flags:
                    flags = " ".join(sorted(line.split(":", 1)[1].split()))
                if model and flags:
                    break
    except OSError:
        pass
    return hashlib.sha256((model + "|" + flags).encode()).hexdigest()[:12]


def build_oracle(force=False):
    deps = [os.path.join(ORACLE_DIR, "evp_oracle.cpp"), os.path.join(ROOT, "include", "evpfft.h"), os.path.join(ORACLE_DIR, "Makefile")]
    marker = "EVPORACLE:" + digest(deps) + ":" + host_signature()
    if force or not has_marker(ORACLE_OUT, marker):
        subprocess.run(["make", "-B", "-C", ORACLE_DIR, f"BUILD_ID={marker}"], check=True, capture_output=True)
    return ORACLE_OUT


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_product(force=force, verbose=True))
    print(build_driver(force=force))
    print(build_oracle(force=force))
