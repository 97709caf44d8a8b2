"""Builds libdensecap_b200.so in-tree with nvcc for sm_100a (no torch dependency, cudart static).

The .so is git-ignored but travels to the GPU box with the repo snapshot.  nvcc cross-compiles
without a GPU, so this runs in the CPU build container too.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libdensecap_b200.so")
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "densecap_b200.h")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMPILE_FLAGS = ARCH_FLAGS + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
LINK_FLAGS = ARCH_FLAGS + ["-shared", "-cudart", "static"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(nvcc, src, obj, defines):
    cmd = [nvcc] + COMPILE_FLAGS + ["-D" + d for d in defines] + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    return src, proc.returncode, proc.stdout + proc.stderr, " ".join(cmd)


def build_library(force=False, verbose=False, defines=(), out_path=None):
    """Compile every .cu under csrc/ (one object per source, in parallel, cached under build/) and link one shared library.
    Returns the library path.  defines / out_path: development variant builds (e.g. -DDC_DISABLE_ROI_TMA to bisect a kernel
    problem on the GPU box)."""
    if out_path is None and not defines and not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libdensecap_b200.so (set NVCC or install the CUDA toolkit)")
    target = out_path or LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    import hashlib
    tag = hashlib.sha1(" ".join(sorted(defines)).encode()).hexdigest()[:10] if defines else "default"
    obj_dir = os.path.join(os.path.dirname(PKG_DIR), "build", "obj_" + tag)
    os.makedirs(obj_dir, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    hdr_time = max(os.path.getmtime(h) for h in headers)
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as pool:
        for src, rc, log, cmd in pool.map(lambda j: _compile_one(nvcc, j[0], j[1], defines), jobs):
            if verbose:
                print(cmd, file=sys.stderr)
            if rc != 0:
                raise RuntimeError("nvcc failed on %s:\n%s" % (src, log))
    tmp = target + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + LINK_FLAGS + ["-o", tmp] + objs
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc link failed:\n%s\n%s" % (proc.stdout, proc.stderr))
    os.replace(tmp, target)
    return target


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
