"""Packed-dense Gaussian exponential family -- torch mirror of svae/distributions/gaussian.py:11-57."""
from __future__ import annotations
import torch
from ..util import T, inv, solve, cholesky


def pack_dense(A, b, *args):
    """gaussian.py:39-53: ((A b 0),(0 c 0),(0 0 d)) of shape (...,N+2,N+2); A may be diagonal (...,N)."""
    lead, N = b.shape[:-1], b.shape[-1]
    z1 = torch.zeros(lead + (N, 1), dtype=b.dtype, device=b.device)
    z2 = torch.zeros(lead + (1, 1), dtype=b.dtype, device=b.device)
    if args:
        c, d = (torch.as_tensor(x, dtype=b.dtype, device=b.device).expand(lead).reshape(lead + (1, 1)) for x in args)
    else:
        c, d = z2, z2
    if A.dim() == b.dim():
        A = A[..., None] * torch.eye(N, dtype=b.dtype, device=b.device)
    b = b[..., None]
    row = lambda *xs: torch.cat(xs, -1)
    return torch.cat((row(A, b, z1), row(T(z1), c, z2), row(T(z1), z2, d)), -2)


def unpack_dense(arr):
    """gaussian.py:55-57."""
    N = arr.shape[-1] - 2
    return arr[..., :N, :N], arr[..., :N, N], arr[..., N, N], arr[..., N + 1, N + 1]


def expectedstats(natparam):
    """gaussian.py:11-17."""
    neghalfJ, h, _, _ = unpack_dense(natparam)
    J = -2 * neghalfJ
    Ex = solve(J, h[..., None])[..., 0]
    ExxT = inv(J) + Ex[..., None] * Ex[..., None, :]
    En = torch.ones(J.shape[:-2], dtype=J.dtype, device=J.device)
    return pack_dense(ExxT, Ex, En, En)


def logZ(natparam):
    """gaussian.py:19-25."""
    neghalfJ, h, a, b = unpack_dense(natparam)
    J = -2 * neghalfJ
    L = cholesky(J)
    return 0.5 * (h * solve(J, h[..., None])[..., 0]).sum() \
        - torch.log(torch.diagonal(L, dim1=-1, dim2=-2)).sum() + (a + b).sum()


def natural_sample(natparam, num_samples, eps=None, generator=None):
    """gaussian.py:27-33: samples (..., S, N); eps (..., N, S) injects the noise."""
    neghalfJ, h, _, _ = unpack_dense(natparam)
    J = -2 * neghalfJ
    L = cholesky(J)
    if eps is None:
        eps = torch.randn(h.shape + (num_samples,), dtype=h.dtype, device=h.device, generator=generator)
    noise = torch.linalg.solve_triangular(T(L), eps, upper=True)
    return solve(J, h[..., None])[..., 0][..., None, :] + T(noise)
