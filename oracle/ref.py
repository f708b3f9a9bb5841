"""Loader + thin drivers for oracle/_ref (the reference's own compiled Cython kernels).

TEST INFRASTRUCTURE ONLY.  Mirrors the composition in the reference's
svae/lds/lds_inference.py:196-202 (cython_natural_lds_inference_general) and the VJP wiring at
svae/lds/lds_inference.py:26-39, one sequence per call exactly like the reference.
"""
from __future__ import annotations
import os, sys, importlib
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")


def available() -> bool:
    try:
        load()
        return True
    except Exception:
        return False


_mods = None


def load():
    """Returns (cython_lds_inference, cython_hmm_inference) modules of the compiled reference."""
    global _mods
    if _mods is None:
        if _REF not in sys.path:
            sys.path.insert(0, _REF)
        lds = importlib.import_module("svae.lds.cython_lds_inference")
        hmm = importlib.import_module("svae.hmm.cython_hmm_inference")
        _mods = (lds, hmm)
    return _mods


def _c(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def lds_estep_one(init_params, pair_params, node_params, num_samples=1, seed=None, want_saved=False):
    """One sequence through the reference's three Cython primitives.

    init_params = (-1/2 J0 (n,n), h0 (n), logZ0); pair_params = (J11, J12, J22, logZ) natural form,
    each (n,n) [homogeneous] or (T-1,n,n); node_params = (J_node (T,n), h_node (T,n), logZ_node (T)).
    The sampler draws its own noise with numpy.random (cython_lds_inference.pyx:333); pass `seed`
    to make it reproducible.  Returns samples (T,S,n), (E_init, E_pair, E_node), lognorm and the
    noise actually used, re-indexed by time: eps_time[t] is the (S,n) draw consumed at time t.
    """
    lds, _ = load()
    init_params = tuple(init_params)
    pair_params = tuple(_c(p) if np.ndim(p) else float(p) for p in pair_params)
    node_params = tuple(_c(p) for p in node_params)
    (msgs, lognorm), saved_f = lds.natural_filter_forward_general(init_params, pair_params, node_params)
    stats, saved_s = lds.natural_smoother_general(msgs, pair_params)
    E_init, E_pair, E_node = stats
    E_pair = list(E_pair)  # lazy map under py3 (pyx:208)
    if seed is not None:
        np.random.seed(seed)
    samples, saved_b = lds.natural_sample_backward(msgs, pair_params, num_samples)
    # saved_b[3] is eps as (N,S,T) F-order with eps[:,:,t] the draw used at time t (pyx:333,345-352)
    eps_time = np.ascontiguousarray(np.asarray(saved_b[3]).T.copy())  # (T,S,N)
    out = dict(samples=np.asarray(samples), E_init=E_init, E_pair=E_pair, E_node=E_node,
               lognorm=float(lognorm), messages=msgs, eps=eps_time)
    if want_saved:
        out["saved"] = (saved_f, saved_s, saved_b)
    return out


def lds_vjp_one(init_params, pair_params, node_params, g_samples, g_stats, g_lognorm, seed=0, num_samples=1):
    """Reference VJP w.r.t. node params (J_node, h_node, logZ_node), composed as autograd would:
    sampler-grad + smoother-grad -> summed message cotangents -> filter-grad
    (svae/lds/lds_inference.py:26-39; pyx:357-409, 236-306, 92-145)."""
    lds, _ = load()
    pair_params = tuple(_c(p) if np.ndim(p) else float(p) for p in pair_params)
    node_params = tuple(_c(p) for p in node_params)
    (msgs, lognorm), saved_f = lds.natural_filter_forward_general(tuple(init_params), pair_params, node_params)
    stats, saved_s = lds.natural_smoother_general(msgs, pair_params)
    np.random.seed(seed)
    samples, saved_b = lds.natural_sample_backward(msgs, pair_params, num_samples)
    eps_time = np.ascontiguousarray(np.asarray(saved_b[3]).T.copy())
    # natural_sample_backward_grad accumulates into its cotangent argument in place
    # (_natural_condition_on_grad writes g_y += ..., cython_gaussian_grads.pxd:524): hand it a copy
    gm_b = lds.natural_sample_backward_grad(np.array(g_samples, dtype=np.float64, order='C'), saved_b)
    # _compute_stats_grad mutates its inputs (pyx:223-224): hand it copies
    g_E_init, g_E_pair, g_E_node = g_stats
    g_copy = (tuple(np.array(x, dtype=np.float64) if np.ndim(x) else x for x in g_E_init),
              tuple(np.array(x, dtype=np.float64) if np.ndim(x) else x for x in g_E_pair),
              tuple(np.array(x, dtype=np.float64) if np.ndim(x) else x for x in g_E_node))
    gm_s = lds.natural_smoother_general_grad(g_copy, saved_s)
    gm = tuple(tuple(a + b for a, b in zip(x, y)) for x, y in zip(gm_b, gm_s))
    g_node = lds.natural_filter_grad((gm, float(g_lognorm)), saved_f)
    return dict(g_J_node=np.asarray(g_node[0]), g_h_node=np.asarray(g_node[1]),
                g_logZ_node=np.asarray(g_node[2]), eps=eps_time, samples=np.asarray(samples),
                lognorm=float(lognorm))


def hmm_logZ(natparam):
    _, hmm = load()
    logZ, saved = hmm.hmm_logZ(tuple(_c(p) for p in natparam))
    return float(logZ), saved


def hmm_estep_slow(natparam):
    """vgrad(hmm_logZ) (svae/hmm/hmm_inference.py:65): logZ and (E[z0], E[trans], E[z_t])."""
    _, hmm = load()
    logZ, saved = hmm.hmm_logZ(tuple(_c(p) for p in natparam))
    g = hmm.hmm_logZ_grad(1.0, saved)
    return float(logZ), tuple(np.asarray(x) for x in g)
