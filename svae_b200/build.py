"""In-tree build of libsvae_b200.so: hand-written sm_100a CUDA behind a C-ABI (include/svae_b200.h).

    python -m svae_b200.build [--force] [-j JOBS] [-v]

nvcc cross-compiles for sm_100a without a GPU.  One object per latent dimension (the kernels are
fully unrolled over n) so the build parallelises; objects are cached in svae_b200/csrc/build/ keyed on
source mtimes.  The .so is git-ignored but NOT gpurun-ignored: it travels to the GPU box.
"""
from __future__ import annotations
import os, subprocess, sys, shutil
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiment hooks: SVAE_BUILD_TAG=<tag> SVAE_BUILD_FLAGS="-DX=1 ..." builds libsvae_b200_<tag>.so
TAG = os.environ.get("SVAE_BUILD_TAG", "")
OBJ = os.path.join(CSRC, "build" + ("_" + TAG if TAG else ""))
INC = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libsvae_b200%s.so" % ("_" + TAG if TAG else ""))
MAX_N = 16
# experiment hook: SVAE_BUILD_NS="10,16" compiles the LDS kernels for those latent dimensions only (stubs for the rest);
# refused for the untagged product library
NS = [int(x) for x in os.environ.get("SVAE_BUILD_NS", "").split(",") if x]
if NS and not TAG:
    raise RuntimeError("SVAE_BUILD_NS needs SVAE_BUILD_TAG: the product library carries every latent dimension")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", INC, "-I", CSRC] + os.environ.get("SVAE_BUILD_FLAGS", "").split()


def _units():
    units = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith("_inst.cu"):
            for n in range(1, MAX_N + 1):
                stub = ["-DSVAE_STUB"] if NS and n not in NS else []
                units.append((f, ["-DSVAE_N=%d" % n] + stub, "%s_n%d.o" % (f[:-8], n)))
        elif f.endswith(".cu"):
            units.append((f, [], f[:-3] + ".o"))
    return units


def _deps_mtime():
    m = os.path.getmtime(__file__)
    for root in (CSRC, INC):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force: bool = False, jobs: int | None = None, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    newest = _deps_mtime()
    units = _units()
    todo = []
    for src, defs, obj in units:
        o = os.path.join(OBJ, obj)
        if force or not os.path.exists(o) or os.path.getmtime(o) < newest:
            todo.append((src, defs, o))
    if not todo and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB

    def cc(u):
        src, defs, o = u
        cmd = [NVCC] + FLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        return u, subprocess.run(cmd, capture_output=True, text=True)

    jobs = jobs or min(max(len(todo), 1), os.cpu_count() or 4)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        for u, r in ex.map(cc, todo):
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s %s:\n%s\n%s" % (u[0], u[1], r.stdout, r.stderr))
            if verbose:
                print(u[2], "\n", r.stderr)
    objs = [os.path.join(OBJ, obj) for _, _, obj in units]
    r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs +
                       ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lcudart"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    j = int(sys.argv[sys.argv.index("-j") + 1]) if "-j" in sys.argv else None
    print(build(force="--force" in sys.argv, jobs=j, verbose="-v" in sys.argv))
