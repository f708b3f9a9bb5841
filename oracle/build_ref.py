"""Build recipe for oracle/_ref: the reference's OWN native E-step kernels.

TEST INFRASTRUCTURE ONLY.  Compiles, unmodified in arithmetic, the reference's Cython sources

    /root/reference/svae/lds/cython_lds_inference.pyx   (filter / smoother / sampler + VJPs)
    /root/reference/svae/hmm/cython_hmm_inference.pyx   (hmm_logZ, hmm_logZ_grad)
    + the .pxd helpers they cimport (cython_gaussian_grads, cython_linalg_grads, cython_util)

from where they lie under /root/reference into extension modules under oracle/_ref/ (git-ignored,
NOT gpurun-ignored, so the built .so travels to the GPU box).  No reference source is copied into
the repository: sources are staged in a throw-away temp dir, and only the compiled .so files (plus
empty package markers) land in oracle/_ref/.

The only edit applied to the staged copy is the Cython-3 compatibility shim described in
SURVEY.md §8c: tuples that mix memoryviews with C scalars are inferred as C structs by Cython 3
("C struct/union member cannot be a memory view"), so each element of those five tuples gets an
explicit <object> coercion -- which is exactly what Cython 0.2x did implicitly.  Sites:
cython_lds_inference.pyx:89,193,194,354 and cython_hmm_inference.pyx:120.  No arithmetic changes.

Usage:  python oracle/build_ref.py            (no-op when /root/reference is absent)
"""
from __future__ import annotations
import os, re, shutil, subprocess, sys, tempfile, glob

REF = os.environ.get("SVAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

SHIM_SITES = {
    "svae/lds/cython_lds_inference.pyx": [89, 193, 194, 354],
    "svae/hmm/cython_hmm_inference.pyx": [120],
}
STAGED = [
    "svae/__init__.py", "svae/cython_util.pxd", "svae/cython_linalg_grads.pxd",
    "svae/lds/__init__.py", "svae/lds/cython_lds_inference.pyx", "svae/lds/cython_gaussian_grads.pxd",
    "svae/hmm/__init__.py", "svae/hmm/cython_hmm_inference.pyx",
]

SETUP = r'''
from setuptools import setup, Extension
import numpy as np
from Cython.Build import cythonize
exts = [Extension('svae.lds.cython_lds_inference', ['svae/lds/cython_lds_inference.pyx'],
                  extra_compile_args=['-O2', '-w']),
        Extension('svae.hmm.cython_hmm_inference', ['svae/hmm/cython_hmm_inference.pyx'],
                  extra_compile_args=['-O2', '-w'])]
setup(ext_modules=cythonize(exts, language_level=2, include_path=['svae', 'svae/lds', '.']),
      include_dirs=[np.get_include()])
'''


def _shim_line(line: str) -> str:
    lhs, rhs = line.split("=", 1)
    rhs = re.sub(r"\b([A-Za-z_][A-Za-z_0-9]*)\b", r"<object>\1", rhs)
    return lhs + "=" + rhs


def available() -> bool:
    return bool(glob.glob(os.path.join(OUT, "svae", "lds", "cython_lds_inference*.so"))) and \
        bool(glob.glob(os.path.join(OUT, "svae", "hmm", "cython_hmm_inference*.so")))


def build(force: bool = False) -> bool:
    if available() and not force:
        return True
    if not os.path.isdir(REF):
        print("[oracle/build_ref] %s absent: keeping whatever is prebuilt in oracle/_ref" % REF)
        return available()
    tmp = tempfile.mkdtemp(prefix="svae_ref_build_")
    try:
        for rel in STAGED:
            dst = os.path.join(tmp, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(REF, rel), dst)
        for rel, lines in SHIM_SITES.items():
            p = os.path.join(tmp, rel)
            src = open(p).read().split("\n")
            for ln in lines:
                src[ln - 1] = _shim_line(src[ln - 1])
            open(p, "w").write("\n".join(src))
        open(os.path.join(tmp, "setup_ref.py"), "w").write(SETUP)
        r = subprocess.run([sys.executable, "setup_ref.py", "build_ext", "--inplace"], cwd=tmp,
                           capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout[-3000:], r.stderr[-3000:])
            return False
        for sub in ("svae", "svae/lds", "svae/hmm"):
            os.makedirs(os.path.join(OUT, sub), exist_ok=True)
            open(os.path.join(OUT, sub, "__init__.py"), "w").close()
        for so in glob.glob(os.path.join(tmp, "svae", "*", "*.so")):
            rel = os.path.relpath(so, tmp)
            shutil.copyfile(so, os.path.join(OUT, rel))
        return available()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref available:", ok)
    sys.exit(0 if ok else 1)
