"""Build libd4gs.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m deblur4dgs_b200.build [--force]

Every translation unit is compiled with
``-gencode arch=compute_100a,code=sm_100a -lineinfo``; ``project.cu``
additionally with ``-fmad=false`` (bit-reproducible projection, see
csrc/project_math.cuh).  The result is ``deblur4dgs_b200/libd4gs.so`` -- it is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libd4gs.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
UNITS = {
    "error.cu": [],
    "project.cu": ["-fmad=false"],
    "binning.cu": [],
    "blend.cu": [],
    "blend_bwd_gp.cu": [],
    "blend_slab_fwd.cu": [],
    "blend_slab_fwd_tc.cu": [],
    "blend_slab_bwd.cu": [],
    "blend_slab_bwd_tc.cu": [],
    "deform.cu": [],
    "combine.cu": [],
    "band_combine.cu": [],
    "correlation.cu": [],
    "camera.cu": [],
    "assemble.cu": [],
}
HEADERS = ["common.cuh", "blend_common.cuh", "slab.cuh", "project_math.cuh", "deform_math.cuh", "camera_math.cuh",
           os.path.join("..", "..", "include", "d4gs.h")]


def _newest_header():
    ts = []
    for h in HEADERS:
        p = os.path.join(CSRC, h)
        if os.path.exists(p):
            ts.append(os.path.getmtime(p))
    return max(ts) if ts else 0.0


def _compile(unit, flags, force):
    src = os.path.join(CSRC, unit)
    obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
    if not os.path.exists(src):
        return None
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src)
            and os.path.getmtime(obj) >= _newest_header()):
        return obj
    cmd = ["nvcc", *ARCH, *COMMON, *flags, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, r.stdout, r.stderr))
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        futs = {u: ex.submit(_compile, u, f, force) for u, f in UNITS.items()}
        objs = [f.result() for f in futs.values()]
    objs = [o for o in objs if o]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = ["nvcc", *ARCH, "-shared", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
