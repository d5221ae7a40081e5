"""Builds the in-tree native library `csrc/libofdg.so` (sm_100a only) with nvcc.

    python optical-flow-2d-data-generation_b200/build.py [--force] [--verbose]

Flags that matter for parity: --fmad=false (device) and -ffp-contract=off (host): float/double
expressions must round like the reference's x86-64 SSE build (no fused multiply-add).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libofdg.so")

SOURCES = ["api.cu", "render.cu", "philox.cu", "warpfields.cu", "host/params.cpp", "host/flatten.cpp", "host/layer.cpp", "host/expand.cpp", "host/texture_io.cpp"]
HEADERS = ["render.cuh", "philox.cuh", "warpfields.cuh", "raster_tile.h", "flat_scene.h", "host/params.hpp", "host/flatten.hpp", "host/affine.hpp",
           "host/mode_tables.inc", "host/layer.hpp", "host/caffe_shim.hpp", "host/expand.hpp", "host/texture_io.hpp"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    deps += [os.path.join(ROOT, "include", "ofdg", h) for h in ("ofdg.h", "scene.h")]
    deps.append(os.path.abspath(__file__))
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "--fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-pthread", "-shared",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB + ".tmp"] + srcs + ["-lz", "-lnvjpeg"]
    cmd += [f for f in os.environ.get("OFDG_NVCC_FLAGS", "").split() if f]  # experiments: -DOFDG_TILE_ROWS=4 ...
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libofdg.so")
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
