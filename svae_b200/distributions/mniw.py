"""Matrix-normal-inverse-Wishart, natural form (A, B, C, d) -- torch mirror of svae/distributions/mniw.py:13-58."""
from __future__ import annotations
import math
import torch
from ..util import symmetrize


def standard_to_natural(nu, S, M, K):
    """mniw.py:25-31."""
    Kinv = torch.linalg.inv(K)
    A = Kinv
    B = Kinv @ M.T
    C = S + M @ B
    return (A, B, C, torch.as_tensor(nu, dtype=S.dtype, device=S.device))


def natural_to_standard(natparam):
    """mniw.py:33-40."""
    A, B, C, d = natparam
    K = symmetrize(torch.linalg.inv(A))
    M = (K @ B).T
    S = C - M @ B
    return d, S, M, K


def expectedstats_standard(nu, S, M, K, fudge=1e-8):
    """mniw.py:42-55: (-1/2 E[A'S^-1 A], E[S^-1 A]', -1/2 E[S^-1], 1/2 E[logdet S^-1]) -- an LDS pair potential."""
    m = M.shape[0]
    eyeS = torch.eye(S.shape[0], dtype=S.dtype, device=S.device)
    eyeK = torch.eye(K.shape[0], dtype=S.dtype, device=S.device)
    E_Sigmainv = nu * symmetrize(torch.linalg.inv(S)) + fudge * eyeS
    E_Sigmainv_A = nu * torch.linalg.solve(S, M)
    E_AT_Sigmainv_A = m * K + nu * symmetrize(M.T @ torch.linalg.solve(S, M)) + fudge * eyeK
    ar = torch.arange(m, dtype=S.dtype, device=S.device)
    E_logdetSigmainv = torch.digamma((nu - ar) / 2.).sum() + m * math.log(2.) - torch.linalg.slogdet(S)[1]
    return (-0.5 * E_AT_Sigmainv_A, E_Sigmainv_A.T, -0.5 * E_Sigmainv, 0.5 * E_logdetSigmainv)


def expectedstats(natparam):
    """mniw.py:19-20."""
    return expectedstats_standard(*natural_to_standard(natparam))


def logZ(natparam):
    """mniw.py:13-17."""
    nu, S, _, K = natural_to_standard(natparam)
    n = S.shape[0]
    return n * nu / 2. * math.log(2.) + torch.special.multigammaln(nu / 2., n) \
        - nu / 2. * torch.linalg.slogdet(S)[1] + n / 2. * torch.linalg.slogdet(K)[1]


def expected_standard_params(natparam):
    """mniw.py:60-64."""
    J11, J12, J22, _ = expectedstats(natparam)
    A = torch.linalg.solve(-2. * J22, J12.T)
    Sigma = torch.linalg.inv(-2 * J22)
    return A, Sigma
