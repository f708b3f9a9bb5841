"""Matrix-normal-inverse-Wishart, natural form (A, B, C, d) -- torch mirror of svae/distributions/mniw.py:13-58."""
from __future__ import annotations
import math
import torch
from ..util import symmetrize, inv, solve, cholesky


_T = lambda X: X.transpose(-1, -2)


def standard_to_natural(nu, S, M, K):
    """mniw.py:25-31 (leading batch dimensions allowed)."""
    Kinv = inv(K)
    A = Kinv
    B = Kinv @ _T(M)
    C = S + M @ B
    return (A, B, C, torch.as_tensor(nu, dtype=S.dtype, device=S.device))


def natural_to_standard(natparam):
    """mniw.py:33-40."""
    A, B, C, d = natparam
    K = symmetrize(inv(A))
    M = _T(K @ B)
    S = C - M @ B
    return d, S, M, K


def expectedstats_standard(nu, S, M, K, fudge=1e-8):
    """mniw.py:42-55: (-1/2 E[A'S^-1 A], E[S^-1 A]', -1/2 E[S^-1], 1/2 E[logdet S^-1]) -- an LDS pair potential.
    Batched over leading dimensions (the K modes of an SLDS in one set of launches)."""
    m = M.shape[-2]
    nu = torch.as_tensor(nu, dtype=S.dtype, device=S.device)
    nub = nu[..., None, None]
    eyeS = torch.eye(S.shape[-1], dtype=S.dtype, device=S.device)
    eyeK = torch.eye(K.shape[-1], dtype=S.dtype, device=S.device)
    SiM = solve(S, M)
    E_Sigmainv = nub * symmetrize(inv(S)) + fudge * eyeS
    E_Sigmainv_A = nub * SiM
    E_AT_Sigmainv_A = m * K + nub * symmetrize(_T(M) @ SiM) + fudge * eyeK
    ar = torch.arange(m, dtype=S.dtype, device=S.device)
    E_logdetSigmainv = torch.digamma((nu[..., None] - ar) / 2.).sum(-1) + m * math.log(2.) - torch.linalg.slogdet(S)[1]
    return (-0.5 * E_AT_Sigmainv_A, _T(E_Sigmainv_A), -0.5 * E_Sigmainv, 0.5 * E_logdetSigmainv)


def expectedstats(natparam):
    """mniw.py:19-20."""
    return expectedstats_standard(*natural_to_standard(natparam))


def logZ(natparam):
    """mniw.py:13-17 (summed over leading batch dimensions, like niw.logZ)."""
    nu, S, _, K = natural_to_standard(natparam)
    n = S.shape[-1]
    return (n * nu / 2. * math.log(2.) + torch.special.multigammaln(nu / 2., n)
            - nu / 2. * torch.linalg.slogdet(S)[1] + n / 2. * torch.linalg.slogdet(K)[1]).sum()


def expected_standard_params(natparam):
    """mniw.py:60-64."""
    J11, J12, J22, _ = expectedstats(natparam)
    A = solve(-2. * J22, _T(J12))
    Sigma = inv(-2 * J22)
    return A, Sigma
