"""sgd / adam over nested-tuple parameters -- the role autograd.optimizers plays in the reference's experiment
scripts (experiments/gmm_svae_synth.py:60: `adam(gradfun, params, num_iters, step_size)`): the optimiser calls
`gradfun(params, i)` (svae_b200.svae.make_gradfun) and steps every leaf.  The PGM block receives its natural
gradient; like the reference, all three blocks go through the same update rule."""
from __future__ import annotations
import torch
from .util import istuple


def _map(f, *structs):
    if istuple(structs[0]):
        return tuple(_map(f, *xs) for xs in zip(*structs))
    return f(*structs)


def sgd(gradfun, params, num_iters=200, step_size=0.1, mass=0.9, callback=None):
    """autograd.optimizers.sgd: momentum SGD.  Returns the final params."""
    vel = _map(torch.zeros_like, params)
    for i in range(num_iters):
        g = gradfun(params, i)
        if callback:
            callback(params, i, g)
        vel = _map(lambda v, gi: mass * v - (1.0 - mass) * gi, vel, g)
        params = _map(lambda p, v: p + step_size * v, params, vel)
    return params


def adam(gradfun, params, num_iters=100, step_size=0.001, b1=0.9, b2=0.999, eps=1e-8, callback=None):
    """autograd.optimizers.adam.  Returns the final params."""
    m = _map(torch.zeros_like, params)
    v = _map(torch.zeros_like, params)
    for i in range(num_iters):
        g = gradfun(params, i)
        if callback:
            callback(params, i, g)
        m = _map(lambda mi, gi: (1 - b1) * gi + b1 * mi, m, g)
        v = _map(lambda vi, gi: (1 - b2) * gi ** 2 + b2 * vi, v, g)
        params = _map(lambda p, mi, vi: p - step_size * (mi / (1 - b1 ** (i + 1))) / (torch.sqrt(vi / (1 - b2 ** (i + 1))) + eps),
                      params, m, v)
    return params
