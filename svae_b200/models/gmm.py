"""GMM-SVAE local mean field -- torch mirror of the reference's svae/models/gmm.py:12-128.

run_inference(prior_natparam, global_natparam, nn_potentials, num_samples) -> (samples, stats, global_kl, local_kl)
local_meanfield(global_natparam, node_potentials) -> (local_stats, prior_stats, natparam, kl)
Block coordinate ascent between the label factor and the Gaussian factor on DETACHED node potentials
(<=100 iterations, |dKL| < 1e-3, gmm.py:90-110), then one differentiable evaluation (gmm.py:73-77).
The per-datapoint work is batched (T, N+2, N+2) dense algebra on the device (N = 2 in the reference's
experiment): global-parameter-sized work that SURVEY.md §2 #12 keeps in PyTorch.  The random initialisation
of the label responsibilities (gmm.py:126-128) can be injected with `label_init` for reproducible parity.
"""
from __future__ import annotations
import sys
import torch
from ..util import unbox, flat, normalize
from ..distributions import dirichlet, categorical, niw, gaussian


def prior_logZ(gmm_natparam):
    d, n = gmm_natparam
    return dirichlet.logZ(d) + niw.logZ(n)


def prior_expectedstats(gmm_natparam):
    d, n = gmm_natparam
    return dirichlet.expectedstats(d), niw.expectedstats(n)


def prior_kl(global_natparam, prior_natparam):
    """gmm.py:54-58."""
    expected_stats = flat(prior_expectedstats(global_natparam))
    natparam_difference = flat(global_natparam) - flat(prior_natparam)
    logZ_difference = prior_logZ(global_natparam) - prior_logZ(prior_natparam)
    return torch.dot(natparam_difference, expected_stats) - logZ_difference


def init_pgm_param(K, N, alpha, niw_conc=10., random_scale=0., generator=None, device=None):
    """gmm.py:33-43."""
    def init_niw_natparam():
        nu, S, kappa = N + niw_conc, (N + niw_conc) * torch.eye(N, dtype=torch.float64), niw_conc
        m = random_scale * torch.randn(N, dtype=torch.float64, generator=generator)
        return niw.standard_to_natural(S, m, kappa, nu)
    dirichlet_natparam = alpha * (torch.rand(K, dtype=torch.float64, generator=generator) if random_scale
                                  else torch.ones(K, dtype=torch.float64))
    niw_natparam = torch.stack([init_niw_natparam() for _ in range(K)])
    return dirichlet_natparam.to(device), niw_natparam.to(device)


def gaussian_meanfield(gaussian_globals, node_potentials, label_stats):
    """gmm.py:112-117."""
    global_potentials = torch.tensordot(label_stats, gaussian_globals, dims=([1], [0]))
    natparam = node_potentials + global_potentials
    stats = gaussian.expectedstats(natparam)
    kl = (node_potentials * stats).sum() - gaussian.logZ(natparam)
    return natparam, stats, kl


def label_meanfield(label_global, gaussian_globals, gaussian_stats):
    """gmm.py:119-124."""
    node_potentials = torch.tensordot(gaussian_stats, gaussian_globals, dims=([1, 2], [1, 2]))
    natparam = node_potentials + label_global
    stats = categorical.expectedstats(natparam)
    kl = (stats * node_potentials).sum() - categorical.logZ(natparam)
    return natparam, stats, kl


def initialize_meanfield(label_global, node_potentials, generator=None):
    """gmm.py:126-128."""
    T, K = node_potentials.shape[0], label_global.shape[0]
    gdev = generator.device if generator is not None else node_potentials.device    # draw where the generator lives
    return normalize(torch.rand(T, K, dtype=torch.float64, generator=generator, device=gdev).to(node_potentials.device))


def meanfield_fixed_point_cuda(label_global, gaussian_globals, node_J, node_h, label_init, tol=1e-3, max_iter=100):
    """The whole coordinate ascent in ONE launch of csrc/gmm_meanfield.cu (one CTA per minibatch).
    node_J, node_h: (T,N) or (B,T,N) diagonal potentials; label_init (T,K) or (B,T,K).  Returns (label_stats, iters, kl)."""
    import ctypes
    from .. import _lib
    dev = node_h.device
    batched = node_h.dim() == 3
    nJ, nh, li = (_lib.f64(x, dev) for x in (node_J, node_h, label_init))
    if not batched:
        nJ, nh, li = nJ[None], nh[None], li[None]
    B, T, N = nh.shape
    K = label_global.shape[0]
    lg, gg = _lib.f64(label_global, dev), _lib.f64(gaussian_globals, dev)
    out = torch.empty((B, T, K), dtype=torch.float64, device=dev)
    iters = torch.zeros(B, dtype=torch.int32, device=dev)
    kl = torch.empty(B, dtype=torch.float64, device=dev)
    P = _lib.ptr
    code = _lib.lib().svae_gmm_meanfield(B, T, K, N, P(nJ.contiguous()), P(nh.contiguous()), P(lg), P(gg), P(li.contiguous()),
                                         float(tol), int(max_iter), P(out), ctypes.c_void_p(iters.data_ptr()), P(kl),
                                         _lib.stream_ptr(dev))
    _lib.check(code, "svae_gmm_meanfield")
    if not batched:
        return out[0], iters[0], kl[0]
    return out, iters, kl


WARN_ITERATION_LIMIT = False   # the reference prints 'iteration limit reached' (gmm.py:107-108); on the CUDA path that needs
                               # the kernel's iteration count on the host, i.e. one synchronisation per E-step: opt-in


def meanfield_fixed_point(label_global, gaussian_globals, node_potentials, tol=1e-3, max_iter=100,
                          label_init=None, generator=None, diagonal=None):
    """gmm.py:90-110.  On a CUDA device with diagonal node precisions the loop runs inside one kernel launch.
    diagonal: True when the caller knows the node precisions are diagonal (local_meanfield does: it packed them from
    (J (T,N), h)); None = find out (costs a device->host synchronisation on CUDA)."""
    kl = float('inf')
    label_stats = initialize_meanfield(label_global, node_potentials, generator) if label_init is None else label_init
    if node_potentials.is_cuda:
        N = node_potentials.shape[-1] - 2
        A = node_potentials[..., :N, :N]
        diag = torch.diagonal(A, dim1=-1, dim2=-2)
        from .. import _lib
        if diagonal is None:
            diagonal = bool((A - torch.diag_embed(diag)).abs().max() == 0)
        if N <= _lib.SVAE_GMM_MAX_N and diagonal:
            labels, iters, _ = meanfield_fixed_point_cuda(label_global, gaussian_globals, diag, node_potentials[..., :N, N],
                                                          label_stats, tol, max_iter)
            if WARN_ITERATION_LIMIT and int(iters.max()) >= max_iter:
                print('iteration limit reached', file=sys.stderr)
            return labels
    for _ in range(max_iter):
        gaussian_natparam, gaussian_stats, gaussian_kl = gaussian_meanfield(gaussian_globals, node_potentials, label_stats)
        label_natparam, label_stats, label_kl = label_meanfield(label_global, gaussian_globals, gaussian_stats)
        # recompute gaussian_kl linear term with new label_stats b/c labels were updated
        gaussian_global_potentials = torch.tensordot(label_stats, gaussian_globals, dims=([1], [0]))
        linear_difference = gaussian_natparam - gaussian_global_potentials - node_potentials
        gaussian_kl = gaussian_kl + (linear_difference * gaussian_stats).sum()
        kl, prev_kl = float(label_kl + gaussian_kl), kl
        if abs(kl - prev_kl) < tol:
            break
    else:
        print('iteration limit reached', file=sys.stderr)
    return label_stats


def local_meanfield(global_natparam, node_potentials, label_init=None, generator=None):
    """gmm.py:62-88."""
    dirichlet_natparam, niw_natparams = global_natparam
    diagonal = node_potentials[0].dim() == node_potentials[1].dim()      # (J (T,N), h (T,N)): the recognition head's output
    node_potentials = gaussian.pack_dense(*node_potentials)
    label_global = dirichlet.expectedstats(dirichlet_natparam)
    gaussian_globals = niw.expectedstats(niw_natparams)
    with torch.no_grad():
        label_stats = meanfield_fixed_point(label_global, gaussian_globals, node_potentials.detach(),
                                            label_init=label_init, generator=generator, diagonal=diagonal or None)
    gaussian_natparam, gaussian_stats, gaussian_kl = gaussian_meanfield(gaussian_globals, node_potentials, label_stats)
    label_natparam, label_stats, label_kl = label_meanfield(label_global, gaussian_globals, gaussian_stats)
    dirichlet_stats = label_stats.sum(0)
    niw_stats = torch.tensordot(label_stats, gaussian_stats, dims=([0], [0]))
    local_stats = label_stats, gaussian_stats
    prior_stats = dirichlet_stats, niw_stats
    natparam = label_natparam, gaussian_natparam
    kl = label_kl + gaussian_kl
    return local_stats, prior_stats, natparam, kl


def run_inference(prior_natparam, global_natparam, nn_potentials, num_samples, label_init=None, eps=None, generator=None):
    """gmm.py:12-16."""
    _, stats, local_natparam, local_kl = local_meanfield(global_natparam, nn_potentials, label_init, generator)
    samples = gaussian.natural_sample(local_natparam[1], num_samples, eps=eps, generator=generator)
    global_kl = prior_kl(global_natparam, prior_natparam)
    return samples, unbox(stats), global_kl, local_kl


def make_encoder_decoder(recognize, decode):
    """gmm.py:18-29."""
    def encode_mean(data, natparam, recogn_params):
        nn_potentials = recognize(recogn_params, data)
        (_, gaussian_stats), _, _, _ = local_meanfield(natparam, nn_potentials)
        _, Ex, _, _ = gaussian.unpack_dense(gaussian_stats)
        return Ex

    def decode_mean(z, phi):
        mu, _ = decode(z, phi)
        return mu.mean(dim=1)
    return encode_mean, decode_mean
