"""LDS-SVAE E-step glue -- torch mirror of the reference's svae/models/lds.py:16-67.

run_inference(prior_natparam, global_natparam, nn_potentials, num_samples)
    -> (samples, global_expected_stats, global_kl, local_kl)                       (lds.py:35-52)
The prior / variational factor on the global parameters is NIW (initial state, packed-dense natural form of
svae/distributions/niw.py) x MNIW (dynamics, (A,B,C,d) of svae/distributions/mniw.py).  [The snapshot's lds.py
imports a tuple-form `svae.lds.niw` that is not in the tree (SURVEY.md ground facts); the in-tree
distributions are used instead -- their expected statistics have exactly the (J, h, logZ1, logZ2) /
(J11, J12, J22, logZ) shapes the LDS routines take.]

nn_potentials = (J_node, h_node) with shape (T,n) [one sequence, like the reference] or (B,T,n) [minibatch of
sequences: statistics are summed over B, KL terms are totals].  The message passing runs in the fused CUDA
kernels (svae_b200/lds/lds_inference.py); gradients reach the node potentials only.
"""
from __future__ import annotations
import torch
from ..util import sub, contract
from ..lds.lds_inference import natural_lds_inference_general
from ..distributions import niw, mniw, gaussian


def lds_prior_expectedstats(natparam):
    """lds.py:23-25."""
    niw_natparam, mniw_natparam = natparam
    return niw.expectedstats(niw_natparam), mniw.expectedstats(mniw_natparam)


def lds_prior_logZ(natparam):
    """lds.py:28-30."""
    niw_natparam, mniw_natparam = natparam
    return niw.logZ(niw_natparam) + mniw.logZ(mniw_natparam)


def lds_prior_kl(global_natparam, prior_natparam, expected_stats=None):
    """lds.py:16-20."""
    if expected_stats is None:
        expected_stats = lds_prior_expectedstats(global_natparam)
    return -contract(sub(prior_natparam, global_natparam), expected_stats) \
        + (lds_prior_logZ(prior_natparam) - lds_prior_logZ(global_natparam))


def run_inference(prior_natparam, global_natparam, nn_potentials, num_samples, eps=None):
    """lds.py:35-52 (the Cython variant: diagonal node potentials)."""
    expected_global = lds_prior_expectedstats(global_natparam)
    init_params = gaussian.unpack_dense(expected_global[0])        # (-1/2 E[J], E[h], logZ terms)
    pair_params = expected_global[1]
    samples, expected_stats, local_normalizer = natural_lds_inference_general(
        (init_params, pair_params), nn_potentials, num_samples, eps=eps)
    (ExxT0, Ex0, n1, n2), (Exx, Exxn, Exnxn, cnt), E_node = expected_stats
    batched = nn_potentials[1].dim() == 3
    red = (lambda x: x.sum(0)) if batched else (lambda x: x)
    # same structure as the prior: packed NIW statistics, MNIW 4-tuple (svae.py:33-34 needs flat() to line up)
    global_expected_stats = (gaussian.pack_dense(red(ExxT0), red(Ex0), red(n1), red(n2)),
                             (red(Exx), red(Exxn), red(Exnxn), red(cnt)))
    local_kl = contract(tuple(nn_potentials[:2]), E_node[:2]) - local_normalizer.sum()
    global_kl = lds_prior_kl(global_natparam, prior_natparam, expected_global)
    return samples, global_expected_stats, global_kl, local_kl


cython_run_inference = run_inference


def run_inference_host(prior_natparam, global_natparam, nn_potentials_host, num_samples, samples_host, chunks=None,
                       generator=None, device=None, local_natparam=None):
    """run_inference (lds.py:35-41) for HOST-resident minibatches: pinned node potentials (J, h) (B,T,n) in, pinned
    samples (B,T,S,n) out, through lds_inference.LdsHostPipeline (pieces of the batch flow H2D copy -> fused kernel ->
    statistics reduction + <node, E_node> contraction -> D2H copy on rotating streams).  Returns the same 4-tuple as
    run_inference with `samples` = samples_host (filled when the current stream has been synchronised), the summed
    global statistics and the KL scalars on the device.  local_natparam = (init_params, pair_params) skips the per-call
    NIW/MNIW maps when the caller already holds them (they are constant across the minibatches of an SVI step)."""
    from ..lds.lds_inference import lds_estep_host, trapezoid_schedule
    B, T, n = nn_potentials_host[1].shape
    expected_global = None
    if local_natparam is None:
        expected_global = lds_prior_expectedstats(global_natparam)
        local_natparam = (gaussian.unpack_dense(expected_global[0]), expected_global[1])
    chunks = chunks if chunks is not None else trapezoid_schedule(B)
    acc, local_kl = lds_estep_host(local_natparam, nn_potentials_host[:2], num_samples, dict(samples=samples_host), chunks=chunks,
                                   generator=generator, device=device, outputs="run_inference")
    nn = n * n
    ExxT0, Ex0 = acc[:nn].reshape(n, n), acc[nn:nn + n]
    o = nn + n
    Exx, Exxn, Exnxn = (acc[o + k * nn:o + (k + 1) * nn].reshape(n, n) for k in range(3))
    Bt = torch.tensor(float(B), dtype=torch.float64, device=acc.device)
    global_expected_stats = (gaussian.pack_dense(ExxT0, Ex0, Bt, Bt), (Exx, Exxn, Exnxn, Bt * (T - 1)))
    global_kl = lds_prior_kl(global_natparam, prior_natparam, expected_global) if prior_natparam is not None else None
    return samples_host, global_expected_stats, global_kl, local_kl


def make_prior_natparam(n, scaling=1., device=None):
    """lds.py:57-67 on the in-tree NIW/MNIW parameterisations."""
    eye = torch.eye(n, dtype=torch.float64, device=device)
    nu, S, mu, kappa = n + 1., 2. * scaling * (n + 1) * eye, torch.zeros(n, dtype=torch.float64, device=device), 1. / (2. * scaling * n)
    M, K = eye.clone(), 1. / (2. * scaling * n) * eye
    return niw.standard_to_natural(S, mu, kappa, nu), mniw.standard_to_natural(nu, S, M, K)
