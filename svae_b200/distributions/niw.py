"""Normal-inverse-Wishart in packed-dense natural form -- torch mirror of svae/distributions/niw.py:15-42."""
from __future__ import annotations
import math
import torch
from ..util import symmetrize, outer, inv, solve, cholesky
from .gaussian import pack_dense, unpack_dense


def natural_to_standard(natparam):
    """niw.py:33-37."""
    A, b, kappa, nu = unpack_dense(natparam)
    m = b / kappa[..., None]
    S = A - outer(b, m)
    return S, m, kappa, nu


def standard_to_natural(S, m, kappa, nu):
    """niw.py:39-42."""
    kappa = torch.as_tensor(kappa, dtype=m.dtype, device=m.device)
    nu = torch.as_tensor(nu, dtype=m.dtype, device=m.device)
    b = kappa[..., None] * m
    A = S + outer(b, m)
    return pack_dense(A, b, kappa, nu)


def expectedstats(natparam, fudge=1e-8):
    """niw.py:15-24."""
    S, m, kappa, nu = natural_to_standard(natparam)
    d = m.shape[-1]
    eye = torch.eye(d, dtype=m.dtype, device=m.device)
    E_J = nu[..., None, None] * symmetrize(inv(S)) + fudge * eye
    E_h = (E_J @ m[..., None])[..., 0]
    E_hTJinvh = d / kappa + (m * E_h).sum(-1)
    ar = torch.arange(d, dtype=m.dtype, device=m.device)
    E_logdetJ = torch.digamma((nu[..., None] - ar) / 2.).sum(-1) + d * math.log(2.) - torch.linalg.slogdet(S)[1]
    return pack_dense(-0.5 * E_J, E_h, -0.5 * E_hTJinvh, 0.5 * E_logdetJ)


def logZ(natparam):
    """niw.py:26-30."""
    S, m, kappa, nu = natural_to_standard(natparam)
    d = m.shape[-1]
    return (d * nu / 2. * math.log(2.) + torch.special.multigammaln(nu / 2., d)
            - nu / 2. * torch.linalg.slogdet(S)[1] - d / 2. * torch.log(kappa)).sum()
