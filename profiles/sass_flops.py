#!/usr/bin/env python
"""Static fp64 instruction count of the constitutive kernels, from the SASS of the built library.

    python profiles/sass_flops.py            # writes profiles/k1_flops.json

For every k_constitutive_p / k_constitutive_t instantiation: DFMA / DADD / DMUL inside the Newton loop (the tightest
backward branch whose body holds at least 150 fp64 instructions) and outside it.  Algorithmic flops per voxel of one launch:
    flops = F0 + F1 * newton_mean,   F = 2 * DFMA + DADD + DMUL
bench.py multiplies by the voxels of the launch and divides by the CUDA-event time -> roofline.achieved (TFLOP/s).
(F0 over-counts a little: both staging paths of the prologue are in the static count, only one executes.)"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lapx_b200 import build  # noqa: E402


def functions(so):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def opname(t):
    parts = t.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    return op.split(".")[0]


def analyse(ins):
    fp = [(a, opname(t)) for a, t in ins if opname(t) in ("DFMA", "DADD", "DMUL")]
    # the Newton loop: the tightest backward branch whose body holds at least 150 fp64 instructions
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.\S+)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and opname(t) == "BRA":
            tgt = int(m.group(1), 16)
            if tgt < a:
                n = sum(1 for x, _ in fp if tgt <= x <= a)
                if n >= 150 and (best is None or a - tgt < best[2] - best[1]):
                    best = (n, tgt, a)
    lo, hi = (best[1], best[2]) if best else (0, -1)
    cnt = {"loop": {"DFMA": 0, "DADD": 0, "DMUL": 0}, "rest": {"DFMA": 0, "DADD": 0, "DMUL": 0}}
    for a, op in fp:
        cnt["loop" if lo <= a <= hi else "rest"][op] += 1
    f = lambda c: 2 * c["DFMA"] + c["DADD"] + c["DMUL"]
    return {"per_newton_iteration": cnt["loop"], "outside_loop": cnt["rest"], "F1": f(cnt["loop"]), "F0": f(cnt["rest"])}


def main():
    so = build.build_product()
    res = {"build_id": "EVPSRC:" + build.source_id(), "kernels": {}}
    for name, ins in functions(so).items():
        if "k_constitutive" not in name:
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"^void evp::|\(.*$", "", dem)
        res["kernels"][short] = analyse(ins)
    path = os.path.join(ROOT, "profiles", "k1_flops.json")
    json.dump(res, open(path, "w"), indent=1, sort_keys=True)
    for k, v in sorted(res["kernels"].items()):
        print(f"{k:55s} F0 {v['F0']:5d}  F1 {v['F1']:5d}")


if __name__ == "__main__":
    main()
