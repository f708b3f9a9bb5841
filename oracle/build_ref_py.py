"""Staging recipe for the reference's own PYTHON layer (GMM / iid / SLDS glue / python LDS twins) under Python 3.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden_py.py to generate fixtures, and by
tests/test_oracle_vs_reference_py.py when /root/reference is present).  Nothing here is imported by the product.

The reference's Python is Python-2 code against a 2016 autograd (SURVEY.md ground facts): it cannot be imported as
it lies.  This recipe copies the files listed in STAGED from /root/reference into a THROW-AWAY temp dir (never into
the repository), applies only MECHANICAL Python-2 -> 3 edits that leave every arithmetic expression untouched:

  * print statement -> function                      (models/gmm.py:108; the matplotlib plotting code after
    gmm.py:130 is cut off, it is not on the path)
  * tuple-parameter lambdas unpacked explicitly      (lds/lds_inference.py:99,100,117-118)
  * implicit relative imports made absolute          (distributions/niw.py:8-9, lds/lds_inference.py:13-24)
  * py2 builtins: xrange, reduce, eager map / zip, itertools.imap, iterator .next, np.object
    (one header inserted after the __future__ line of every staged module)

and puts two shim packages next to them:

  * autograd -> numpy / scipy (autograd.numpy IS numpy here: there is no tape, values are plain arrays; getval is
    the identity; primitive / primitive_with_aux return the undifferentiated function, dropping the aux output
    exactly as autograd@0f026ab's primitive_with_aux does on the forward pass)
  * toolz.curry

(and one numpy-1.x rule, linalg.solve with a stack of vectors, see the shim source)

The compiled Cython kernels of oracle/_ref (oracle/build_ref.py) are linked into the staged package so that
svae.lds.lds_inference and svae.hmm import them exactly as the reference does.  pyhsmm (hmm_estep's backend,
svae/hmm/hmm_inference.py:8-10) is absent: hmm_inference.py is NOT staged; callers get hmm_estep backed by the
reference's in-tree equivalent vgrad(hmm_logZ) (hmm_inference.py:65) through the compiled hmm_logZ / hmm_logZ_grad.

models/slds_svae.py and models/lds.py import modules that do not exist in the tree (svae.lds.niw, svae.lds.mniw,
svae.hmm.dirichlet, lds_svae): they cannot be imported even under Python 2.  Their self-contained functions are
exec'd from the staged source by LINE RANGE into a namespace bound to the staged svae.util / distributions
(slds_functions()).

Usage:
    import oracle.build_ref_py as R
    with R.staged() as ref:          # ref.gmm, ref.iid, ref.util, ref.lds_inference, ref.slds (namespace), ...
        ...
"""
from __future__ import annotations
import contextlib, glob, importlib, os, re, shutil, sys, tempfile, types

REF = os.environ.get("SVAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
CYREF = os.path.join(HERE, "_ref")

STAGED = [
    "svae/__init__.py", "svae/util.py",
    "svae/distributions/__init__.py", "svae/distributions/gaussian.py", "svae/distributions/niw.py",
    "svae/distributions/mniw.py", "svae/distributions/dirichlet.py", "svae/distributions/categorical.py",
    "svae/models/__init__.py", "svae/models/gmm.py", "svae/models/iid_gausisan.py", "svae/models/slds_svae.py",
    "svae/lds/__init__.py", "svae/lds/gaussian.py", "svae/lds/gaussian_nochol.py", "svae/lds/lds_inference.py",
    "svae/hmm/__init__.py",
]
NOT_IMPORTABLE = {"svae/models/slds_svae.py"}      # exec'd by line range only
TRUNCATE_AT = {"svae/models/gmm.py": "### plotting util for 2D"}   # matplotlib plotting code (py2 raise syntax): out of scope

HEADER = ("from functools import reduce\nimport builtins as _b\n"
          "map = lambda *a: list(_b.map(*a))\nzip = lambda *a: list(_b.zip(*a))\nxrange = range\n")

# exact-text line edits: (file, line number, expected original (stripped), replacement (same indentation))
LINE_EDITS = [
    ("svae/models/gmm.py", 108, "print 'iteration limit reached'", "print('iteration limit reached')"),
    ("svae/lds/lds_inference.py", 99,
     "condition = lambda node_param: lambda (J, h): natural_condition_on_general(J, h, *node_param)",
     "condition = lambda node_param: lambda J_h: natural_condition_on_general(J_h[0], J_h[1], *node_param)"),
    ("svae/lds/lds_inference.py", 100,
     "predict = lambda pair_param: lambda (J, h): natural_predict(J, h, *pair_param)",
     "predict = lambda pair_param: lambda J_h: natural_predict(J_h[0], J_h[1], *pair_param)"),
    ("svae/lds/lds_inference.py", 117,
     "sample = lambda (J11, J12), (J_filt, h_filt): lambda next_sample: \\",
     "sample = lambda J11_J12, Jf_hf: lambda next_sample: \\"),
    ("svae/lds/lds_inference.py", 118,
     "natural_sample(*natural_condition_on(J_filt, h_filt, next_sample, J11, J12))",
     "natural_sample(*natural_condition_on(Jf_hf[0], Jf_hf[1], next_sample, J11_J12[0], J11_J12[1]))"),
    ("svae/util.py", 93, "nexts = cycle(iter(it).next for it in iterables)",
     "nexts = cycle(iter(it).__next__ for it in iterables)"),
    ("svae/util.py", 172, "isobjarray = lambda x: isinstance(x, np.ndarray) and x.dtype == np.object",
     "isobjarray = lambda x: isinstance(x, np.ndarray) and x.dtype == object"),
]
# regex edits (imports only)
TEXT_EDITS = {
    "svae/util.py": [(r"from itertools import islice, imap, cycle", "from itertools import islice, cycle\nimap = map")],
    "svae/distributions/niw.py": [(r"^from gaussian import", "from svae.distributions.gaussian import"),
                                  (r"^import mniw\b", "from svae.distributions import mniw")],
    "svae/lds/lds_inference.py": [(r"^from gaussian import", "from svae.lds.gaussian import"),
                                  (r"^from gaussian_nochol import", "from svae.lds.gaussian_nochol import"),
                                  (r"^from cython_lds_inference import", "from svae.lds.cython_lds_inference import")],
}

AUTOGRAD_SHIM = r'''
"""autograd -> numpy/scipy shim (no tape): see oracle/build_ref_py.py."""
import sys, types
import numpy as _numpy, numpy.random, numpy.linalg
import scipy as _scipy, scipy.linalg, scipy.special

def _mod(name, **attrs):
    m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m; return m

# autograd.numpy is numpy -- with ONE numpy-1.x rule restored: linalg.solve(a, b) treated b as a stack of VECTORS when
# b.ndim == a.ndim - 1 (numpy 2 does so only for b.ndim == 1); the reference (2016) relies on it, e.g.
# svae/distributions/gaussian.py:14 solve(J (T,N,N), h (T,N)).
def _solve_np1(a, b):
    a, b = _numpy.asarray(a), _numpy.asarray(b)
    if b.ndim == a.ndim - 1 and b.ndim > 1:
        return _numpy.linalg.solve(a, b[..., None])[..., 0]
    return _numpy.linalg.solve(a, b)
_linalg = _mod("autograd.numpy.linalg", __getattr__=lambda name: getattr(_numpy.linalg, name), solve=_solve_np1)
_np = _mod("autograd.numpy", __getattr__=lambda name: getattr(_numpy, name), linalg=_linalg, random=_numpy.random)
sys.modules["autograd.numpy.random"] = _numpy.random
_sp = _mod("autograd.scipy", linalg=_scipy.linalg, special=_scipy.special)
sys.modules["autograd.scipy.linalg"] = _scipy.linalg
sys.modules["autograd.scipy.special"] = _scipy.special
_sp.misc = _mod("autograd.scipy.misc", logsumexp=_scipy.special.logsumexp)

def _no_tape(*a, **k):
    raise NotImplementedError("autograd shim: no differentiation (values only)")
grad = value_and_grad = lambda f, *a, **k: _no_tape

def _flatten(value):
    """autograd.util.flatten: (1-D vector, unflatten)."""
    if isinstance(value, (tuple, list)):
        parts = [_flatten(v) for v in value]
        vec = _numpy.concatenate([p[0] for p in parts]) if parts else _numpy.zeros(0)
        sizes = [p[0].size for p in parts]
        def unflatten(v, parts=parts, sizes=sizes, kind=type(value)):
            out, o = [], 0
            for (pv, un), s in zip(parts, sizes):
                out.append(un(v[o:o + s])); o += s
            return kind(out) if kind is list else tuple(out)
        return vec, unflatten
    arr = _numpy.asarray(value, dtype=float)
    return arr.ravel(), (lambda v, shape=arr.shape: _numpy.reshape(v, shape))

_mod("autograd.util", flatten=_flatten, make_tuple=lambda *a: tuple(a))

class _Prim(object):
    def __init__(self, f, aux=False): self.f, self.aux = f, aux
    def __call__(self, *a, **k):
        out = self.f(*a, **k)
        return out[0] if self.aux else out
    def defgrad(self, *a, **k): pass
    def defgrads(self, *a, **k): pass
_mod("autograd.core", getval=lambda x: x, primitive=lambda f: _Prim(f), primitive_with_aux=lambda f: _Prim(f, True))
class TupleNode(tuple): pass
class ListNode(list): pass
_mod("autograd.container_types", TupleNode=TupleNode, ListNode=ListNode)
_mod("autograd.convenience_wrappers", grad_and_aux=lambda f, *a, **k: _no_tape, value_and_grad=value_and_grad)
numpy, scipy = _np, _sp    # `import autograd.numpy as np` binds the package attribute
'''

TOOLZ_SHIM = r'''
import functools, inspect
def curry(f):
    n = len(inspect.signature(f).parameters)
    @functools.wraps(f)
    def w(*a, **k):
        if len(a) + len(k) >= n:
            return f(*a, **k)
        return functools.partial(w, *a, **k)
    return w
'''


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "svae"))


def _transform(rel: str, text: str) -> str:
    if rel in TRUNCATE_AT:
        text = text[:text.index(TRUNCATE_AT[rel])]
    lines = text.split("\n")
    for f, ln, want, repl in LINE_EDITS:
        if f != rel:
            continue
        got = lines[ln - 1]
        if got.strip() != want:
            raise RuntimeError("%s:%d drifted: %r" % (rel, ln, got))
        lines[ln - 1] = got[:len(got) - len(got.lstrip())] + repl
    text = "\n".join(lines)
    for pat, repl in TEXT_EDITS.get(rel, []):
        text, n = re.subn(pat, repl, text, flags=re.M)
        if n == 0:
            raise RuntimeError("%s: pattern %r not found" % (rel, pat))
    # header after the __future__ import (or at the top)
    m = re.search(r"^from __future__ import [^\n]*\n", text, flags=re.M)
    pos = m.end() if m else 0
    return text[:pos] + HEADER + text[pos:]


def stage(tmp: str) -> None:
    for rel in STAGED:
        dst = os.path.join(tmp, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        src = open(os.path.join(REF, rel)).read()
        if rel in NOT_IMPORTABLE:
            dst += ".txt"          # kept as text: exec'd by line range
            open(dst, "w").write(src)
        else:
            open(dst, "w").write(_transform(rel, src) if rel.endswith(".py") and not rel.endswith("__init__.py") else src)
    os.makedirs(os.path.join(tmp, "autograd"), exist_ok=True)
    open(os.path.join(tmp, "autograd", "__init__.py"), "w").write(AUTOGRAD_SHIM)
    os.makedirs(os.path.join(tmp, "toolz"), exist_ok=True)
    open(os.path.join(tmp, "toolz", "__init__.py"), "w").write(TOOLZ_SHIM)
    # the reference's compiled Cython kernels, where the reference's own imports expect them
    for so in glob.glob(os.path.join(CYREF, "svae", "*", "*.so")):
        rel = os.path.relpath(so, CYREF)
        os.symlink(so, os.path.join(tmp, rel))


def _exec_ranges(ns: dict, path: str, ranges) -> None:
    src = open(path).read().split("\n")
    for a, b in ranges:
        code = "\n".join(src[a - 1:b])
        exec(compile("\n" * (a - 1) + code, path, "exec"), ns)


def slds_functions(ref) -> types.SimpleNamespace:
    """The self-contained functions of svae/models/slds_svae.py, exec'd by line range:
    lds_meanfield :80-83, get_all_lds_local_natparams :86-89, get_var_lds_local_natparam :92-103,
    hmm_meanfield :108-115, get_arhmm_local_nodeparams :131-147, get_hmm_vlb :150-154,
    optimize_local_meanfield :159-175, get_global_stats :229-243.
    Free names are bound to the staged reference modules; the four that the stale imports would have supplied are
    injected by the caller through the returned namespace (lds_prior_expectedstats, hmm_estep, hmm_logZ,
    get_hmm_local_natparam, initialize_local_meanfield)."""
    import numpy as np, sys as _sys
    u = ref.util
    ns = dict(np=np, npr=np.random, sys=_sys, make_tuple=lambda *a: tuple(a), contract=u.contract, scale=u.scale, unbox=u.unbox,
              sub=u.sub, add=u.add, map=lambda *a: list(map(*a)), zip=lambda *a: list(zip(*a)), xrange=range,
              print_function=None, tuplify=u.tuplify, shape=u.shape, randn_like=u.randn_like, partial=__import__('functools').partial,
              )
    # slds_svae.py:13-16 binds these three names to the CYTHON drivers (lds_inference.py:196-202,232-237,260-264); the
    # reference's own canonicalisers (lds_inference.py:62-82) are applied first, as its python drivers do, so that the
    # 4-tuple init parameters the SLDS builds and (J, h) node potentials reach the kernels in the shape they read
    li = ref.lds_inference
    canon = lambda natparam, node: ((li._canonical_init_params(natparam[0]), natparam[1]), li._canonical_node_params(tuple(node)))
    ns["natural_lds_estep_general"] = lambda natparam, node: li.cython_natural_lds_estep_general(*canon(natparam, node))
    ns["natural_lds_inference_general"] = \
        lambda natparam, node, num_samples=1: li.cython_natural_lds_inference_general(*canon(natparam, node), num_samples=num_samples)
    ns["natural_lds_sample"] = \
        lambda natparam, node, num_samples=1: li.cython_natural_lds_sample(*canon(natparam, node), num_samples=num_samples)
    _exec_ranges(ns, ref.slds_source, [(80, 83), (86, 89), (92, 103), (108, 115), (131, 147), (150, 154), (159, 175),
                                       (229, 243)])
    return types.SimpleNamespace(**{k: v for k, v in ns.items() if callable(v) or k == "np"}, _ns=ns)


@contextlib.contextmanager
def staged():
    """Stage, import, yield a namespace of reference modules, and clean up (sys.path / sys.modules restored)."""
    if not available():
        raise RuntimeError("%s is absent" % REF)
    tmp = tempfile.mkdtemp(prefix="svae_ref_py_")
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k == "svae" or k.startswith("svae.") or k == "autograd"
                  or k.startswith("autograd.") or k == "toolz"}
    for k in saved_mods:
        del sys.modules[k]
    try:
        stage(tmp)
        sys.path.insert(0, tmp)
        importlib.import_module("autograd")
        ref = types.SimpleNamespace(
            root=tmp,
            util=importlib.import_module("svae.util"),
            gaussian=importlib.import_module("svae.distributions.gaussian"),
            niw=importlib.import_module("svae.distributions.niw"),
            mniw=importlib.import_module("svae.distributions.mniw"),
            dirichlet=importlib.import_module("svae.distributions.dirichlet"),
            categorical=importlib.import_module("svae.distributions.categorical"),
            gmm=importlib.import_module("svae.models.gmm"),
            iid=importlib.import_module("svae.models.iid_gausisan"),
            lds_gaussian=importlib.import_module("svae.lds.gaussian"),
            lds_inference=importlib.import_module("svae.lds.lds_inference"),
            cython_hmm=importlib.import_module("svae.hmm.cython_hmm_inference"),
            slds_source=os.path.join(tmp, "svae/models/slds_svae.py.txt"))
        ref.slds = slds_functions(ref)
        yield ref
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "svae" or k.startswith("svae.") or k == "autograd"
                  or k.startswith("autograd.") or k == "toolz"]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    with staged() as ref:
        print("staged reference python:", sorted(k for k in vars(ref) if not k.startswith("_")))
