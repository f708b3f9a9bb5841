"""fp64 NumPy restatement of the reference's GMM local mean field (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/svae/models/gmm.py:62-128 with /root/reference/svae/distributions/gaussian.py:11-57,
categorical.py:6-9, dirichlet.py:5-11, niw.py:15-42.  The label-responsibility initialisation
(gmm.py:126-128, npr.rand) is injected.  Parity status: the reference's Python layer cannot be imported
under Python 3 (SURVEY.md ground facts) and its tests hold no vectors for this path -- PARITY UNPINNED
beyond the distribution identities the reference's own tests check (tests/test_distributions.py).
"""
import numpy as np
from scipy.special import digamma, logsumexp


def pack_dense(A, b, c=None, d=None):
    lead, N = b.shape[:-1], b.shape[-1]
    out = np.zeros(lead + (N + 2, N + 2))
    out[..., :N, :N] = A[..., None] * np.eye(N) if A.ndim == b.ndim else A
    out[..., :N, N] = b
    if c is not None:
        out[..., N, N] = c
        out[..., N + 1, N + 1] = d
    return out


def unpack_dense(arr):
    N = arr.shape[-1] - 2
    return arr[..., :N, :N], arr[..., :N, N], arr[..., N, N], arr[..., N + 1, N + 1]


def gaussian_expectedstats(natparam):
    neghalfJ, h, _, _ = unpack_dense(natparam)
    J = -2 * neghalfJ
    Ex = np.linalg.solve(J, h[..., None])[..., 0]
    ExxT = np.linalg.inv(J) + Ex[..., None] * Ex[..., None, :]
    En = np.ones(J.shape[:-2])
    return pack_dense(ExxT, Ex, En, En)


def gaussian_logZ(natparam):
    neghalfJ, h, a, b = unpack_dense(natparam)
    J = -2 * neghalfJ
    L = np.linalg.cholesky(J)
    return 0.5 * np.sum(h * np.linalg.solve(J, h[..., None])[..., 0]) \
        - np.sum(np.log(np.diagonal(L, axis1=-1, axis2=-2))) + np.sum(a + b)


def niw_expectedstats(natparam, fudge=1e-8):
    A, b, kappa, nu = unpack_dense(natparam)
    m = b / kappa[..., None]
    S = A - b[..., :, None] * m[..., None, :]
    d = m.shape[-1]
    Sinv = np.linalg.inv(S)
    E_J = nu[..., None, None] * (Sinv + np.swapaxes(Sinv, -1, -2)) / 2 + fudge * np.eye(d)
    E_h = (E_J @ m[..., None])[..., 0]
    E_hTJinvh = d / kappa + np.sum(m * E_h, -1)
    E_logdetJ = np.sum(digamma((nu[..., None] - np.arange(d)) / 2.), -1) + d * np.log(2.) - np.linalg.slogdet(S)[1]
    return pack_dense(-0.5 * E_J, E_h, -0.5 * E_hTJinvh, 0.5 * E_logdetJ)


def dirichlet_expectedstats(natparam):
    alpha = natparam + 1
    return digamma(alpha) - digamma(alpha.sum(-1, keepdims=True))


def gaussian_meanfield(gaussian_globals, node_potentials, label_stats):
    natparam = node_potentials + np.tensordot(label_stats, gaussian_globals, [1, 0])
    stats = gaussian_expectedstats(natparam)
    return natparam, stats, np.sum(node_potentials * stats) - gaussian_logZ(natparam)


def label_meanfield(label_global, gaussian_globals, gaussian_stats):
    node_potentials = np.tensordot(gaussian_stats, gaussian_globals, [[1, 2], [1, 2]])
    natparam = node_potentials + label_global
    stats = np.exp(natparam - logsumexp(natparam, axis=-1, keepdims=True))
    return natparam, stats, np.sum(stats * node_potentials) - np.sum(logsumexp(natparam, axis=-1))


def local_meanfield(global_natparam, node_potentials, label_init, tol=1e-3, max_iter=100):
    dirichlet_natparam, niw_natparams = global_natparam
    node_potentials = pack_dense(*node_potentials)
    label_global = dirichlet_expectedstats(dirichlet_natparam)
    gaussian_globals = niw_expectedstats(niw_natparams)
    kl, label_stats, iters = np.inf, label_init, 0
    for iters in range(1, max_iter + 1):
        gaussian_natparam, gaussian_stats, gaussian_kl = gaussian_meanfield(gaussian_globals, node_potentials, label_stats)
        label_natparam, label_stats, label_kl = label_meanfield(label_global, gaussian_globals, gaussian_stats)
        lin = gaussian_natparam - np.tensordot(label_stats, gaussian_globals, [1, 0]) - node_potentials
        gaussian_kl = gaussian_kl + np.sum(lin * gaussian_stats)
        kl, prev_kl = label_kl + gaussian_kl, kl
        if abs(kl - prev_kl) < tol:
            break
    gaussian_natparam, gaussian_stats, gaussian_kl = gaussian_meanfield(gaussian_globals, node_potentials, label_stats)
    label_natparam, label_stats, label_kl = label_meanfield(label_global, gaussian_globals, gaussian_stats)
    prior_stats = label_stats.sum(0), np.tensordot(label_stats, gaussian_stats, [0, 0])
    return (label_stats, gaussian_stats), prior_stats, (label_natparam, gaussian_natparam), label_kl + gaussian_kl, iters
