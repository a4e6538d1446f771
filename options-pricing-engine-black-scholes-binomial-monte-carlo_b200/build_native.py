#!/usr/bin/env python3
"""Build liboptp_b200.so (hand-written sm_100a CUDA behind the C ABI of include/optp_b200.h).

    python build_native.py [--force] [--verbose]

Every translation unit is compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
(mc_fp64.cu additionally with -fmad=false: it restates the reference's operation order and
must not contract a*b+c), then linked into optpricer_b200/_lib/liboptp_b200.so.  The library
is built in-tree so it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "optpricer_b200", "_lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(OUT_DIR, "liboptp_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "--extended-lambda", "-Xcompiler", "-fPIC",
          "-Xptxas", "-v"]
UNITS = {
    "optp_api.cu": [],
    "mc_host.cu": [],
    "mc_fp32.cu": [],
    "mc_fp64.cu": ["-fmad=false"],
    "crr_batch.cu": [],
    "dupire_table.cu": ["-fmad=false"],
    "microbench.cu": [],
    "rng_check.cu": [],
    "comm.cu": [],
}
HEADERS = ["optp_common.cuh", "mc_kernels.cuh", os.path.join("..", "..", "include", "optp_b200.h")]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(nvcc, unit, extra, verbose):
    src = os.path.join(CSRC, unit)
    obj = os.path.join(OBJ_DIR, unit.replace(".cu", ".o"))
    # OPTP_NVCC_FLAGS: extra flags for experiments (e.g. -DOPTP_ZONE_MINB4=5); use with --force
    cmd = [nvcc, *ARCH, *COMMON, *extra, *os.environ.get("OPTP_NVCC_FLAGS", "").split(), "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {unit}:\n{log[-4000:]}")
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    todo = []
    for unit, extra in UNITS.items():
        obj = os.path.join(OBJ_DIR, unit.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, unit), *hdrs]):
            todo.append((unit, extra))
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        list(ex.map(lambda ue: _compile(nvcc, ue[0], ue[1], verbose), todo))
    objs = [os.path.join(OBJ_DIR, u.replace(".cu", ".o")) for u in UNITS]
    if todo or force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
