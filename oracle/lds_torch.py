"""torch.float64 CPU twin of oracle/lds_numpy.py (TEST INFRASTRUCTURE ONLY).

Same algorithm, same reference citations (svae/lds/gaussian.py:37-121, svae/lds/lds_inference.py:86-202),
written with differentiable torch ops so that torch.autograd supplies the VJP oracle for the hand-written
CUDA backward kernels -- the role autograd + the Cython *_grad functions play in the reference
(lds_inference.py:26-39).  Pinned against the reference's own VJPs through tests/golden/*.npz.
"""
from __future__ import annotations
import torch

_T = lambda x: x.transpose(-1, -2)


def _solve_tri(L, b, trans=False):
    vec = b.dim() == L.dim() - 1
    bb = b.unsqueeze(-1) if vec else b
    out = torch.linalg.solve_triangular(_T(L) if trans else L, bb, upper=trans)
    return out.squeeze(-1) if vec else out


def natural_predict(J, h, J11, J12, J22, logZ):
    """gaussian.py:52-66."""
    J, J11, J12, J22 = -2 * J, -2 * J11, -J12, -2 * J22
    L = torch.linalg.cholesky(J + J11)
    v = _solve_tri(L, h)
    lognorm = 0.5 * (v * v).sum(-1) - torch.log(torch.diagonal(L, dim1=-1, dim2=-2)).sum(-1)
    v2 = _solve_tri(L, v, trans=True)
    h_predict = -torch.einsum('...ji,...j->...i', J12.expand(L.shape), v2)
    temp = _solve_tri(L, J12.expand(L.shape))
    J_predict = J22 - _T(temp) @ temp
    return (-0.5 * J_predict, h_predict), lognorm + logZ


def natural_lognorm(J, h):
    """gaussian.py:80-84."""
    L = torch.linalg.cholesky(-2 * J)
    v = _solve_tri(L, h)
    return 0.5 * (v * v).sum(-1) - torch.log(torch.diagonal(L, dim1=-1, dim2=-2)).sum(-1)


def natural_rts_backward_step(next_smooth, next_pred, filtered, pair_param):
    """gaussian.py:87-106."""
    (Jns, hns, mun), (Jnp, hnp), (Jf, hf) = next_smooth, next_pred, filtered
    J11, J12, J22 = pair_param[:3]
    Jns, Jnp, Jf, J11, J12, J22 = -2 * Jns, -2 * Jnp, -2 * Jf, -2 * J11, -J12, -2 * J22
    J11, J22 = Jf + J11, Jns - Jnp + J22
    L = torch.linalg.cholesky(J22)
    J12b = J12.expand(L.shape)
    temp = _solve_tri(L, _T(J12b))
    Js = J11 - _T(temp) @ temp
    hs = hf - torch.einsum('...ji,...j->...i', temp, _solve_tri(L, hns - hnp))
    sigma = torch.linalg.inv(Js)
    mu = torch.einsum('...ij,...j->...i', sigma, hs)
    ExnxT = -_solve_tri(L, _solve_tri(L, _T(J12b) @ sigma), trans=True) + mun[..., :, None] * mu[..., None, :]
    ExxT = sigma + mu[..., :, None] * mu[..., None, :]
    return -0.5 * Js, hs, (mu, ExxT, ExnxT)


def natural_sample(J, h, eps):
    """gaussian.py:69-77, noise injected.  J (B,n,n), h (B,S,n), eps (B,S,n)."""
    L = torch.linalg.cholesky(-2 * J)
    noise = _solve_tri(L, _T(eps), trans=True)
    mean = _solve_tri(L, _solve_tri(L, _T(h)), trans=True)
    return _T(mean + noise)


def _pair_at(pair_params, t):
    J11, J12, J22, logZ = pair_params
    if J11.dim() == 2:
        return J11, J12, J22, logZ
    return J11[..., t, :, :], J12[..., t, :, :], J22[..., t, :, :], logZ[..., t]


def natural_lds_inference_general(natparam, node_params, eps):
    """lds_inference.py:185-202, batched; node J diagonal (B,T,n) or dense (B,T,n,n) (gaussian.py:46-49; the
    dense case returns the full E[x x'] as E_node_stats[0], lds_inference.py:159-167).  Everything differentiable."""
    (J0, h0, logZ0), pair_params = natparam
    Jn, hn = node_params[0], node_params[1]
    logZn = node_params[2] if len(node_params) > 2 else torch.zeros(hn.shape[:2], dtype=hn.dtype)
    B, T, n = hn.shape
    eye = torch.eye(n, dtype=hn.dtype)
    J, h = J0.expand(B, n, n), h0.expand(B, n)
    lognorm = torch.zeros(B, dtype=hn.dtype) + logZ0
    Jp, hp, Jf, hf = [], [], [], []
    for t in range(T):                                            # lds_inference.py:86-106
        Jp.append(J); hp.append(h)
        J, h = J + (Jn[:, t] if Jn.dim() == 4 else Jn[:, t, :, None] * eye), h + hn[:, t]
        lognorm = lognorm + logZn[:, t]
        Jf.append(J); hf.append(h)
        if t < T - 1:
            (J, h), term = natural_predict(J, h, *_pair_at(pair_params, t))
            lognorm = lognorm + term
    lognorm = lognorm + natural_lognorm(J, h)
    # smoother                                                       lds_inference.py:127-178
    sigma = torch.linalg.inv(-2 * Jf[-1])
    mu = torch.einsum('bij,bj->bi', sigma, hf[-1])
    Js, hs = Jf[-1], hf[-1]
    stats = [(mu, sigma + mu[:, :, None] * mu[:, None, :], None)]
    for t in range(T - 2, -1, -1):
        Js, hs, (mu_t, ExxT, ExnxT) = natural_rts_backward_step(
            (Js, hs, mu), (Jp[t + 1], hp[t + 1]), (Jf[t], hf[t]), _pair_at(pair_params, t))
        mu = mu_t
        stats.insert(0, (mu, ExxT, ExnxT))
    Ex = torch.stack([s[0] for s in stats], 1)
    ExxT = torch.stack([s[1] for s in stats], 1)
    E_init = (ExxT[:, 0], Ex[:, 0])
    if T > 1:
        ExxnT = torch.stack([_T(s[2]) for s in stats[:-1]], 1)
        E_pair = (ExxT[:, :-1], ExxnT, ExxT[:, 1:])
        if pair_params[0].dim() == 2:
            E_pair = tuple(x.sum(1) for x in E_pair)
    else:
        z = torch.zeros(B, n, n, dtype=hn.dtype)
        E_pair = (z, z, z)
    E_node = (ExxT if Jn.dim() == 4 else torch.diagonal(ExxT, dim1=-1, dim2=-2), Ex)
    # sampler                                                        lds_inference.py:109-124
    S = eps.shape[2]
    x = natural_sample(Jf[-1], hf[-1][:, None, :].expand(B, S, n), eps[:, -1])
    xs = [x]
    for t in range(T - 2, -1, -1):
        J11, J12 = _pair_at(pair_params, t)[:2]
        hc = hf[t][:, None, :] + torch.einsum('...ij,bsj->bsi', J12, x)
        x = natural_sample(Jf[t] + J11, hc, eps[:, t])
        xs.insert(0, x)
    samples = torch.stack(xs, 1)
    return samples, (E_init, E_pair, E_node), lognorm


def vjp_node_params(natparam, node_params, eps, g_samples, g_stats, g_lognorm):
    """Cotangents of (J_node, h_node) given cotangents of every E-step output.

    g_stats = ((g_E_init_xx, g_E_init_x), (g_E_pair_xx, g_E_pair_xxn, g_E_pair_xnxn), (g_E_node_xx, g_E_node_x)).
    """
    t64 = lambda x: torch.as_tensor(x, dtype=torch.float64)
    (J0, h0, logZ0), pair = natparam
    natparam = ((t64(J0), t64(h0), t64(logZ0)), tuple(t64(x) for x in pair))
    Jn = t64(node_params[0]).clone().requires_grad_(True)
    hn = t64(node_params[1]).clone().requires_grad_(True)
    node = (Jn, hn) + ((t64(node_params[2]),) if len(node_params) > 2 else ())
    samples, (E_init, E_pair, E_node), lognorm = natural_lds_inference_general(natparam, node, t64(eps))
    (gi_xx, gi_x), (gp_xx, gp_xxn, gp_xnxn), (gn_xx, gn_x) = g_stats
    total = (samples * t64(g_samples)).sum() + (lognorm * t64(g_lognorm)).sum() \
        + (E_init[0] * t64(gi_xx)).sum() + (E_init[1] * t64(gi_x)).sum() \
        + (E_pair[0] * t64(gp_xx)).sum() + (E_pair[1] * t64(gp_xxn)).sum() + (E_pair[2] * t64(gp_xnxn)).sum() \
        + (E_node[0] * t64(gn_xx)).sum() + (E_node[1] * t64(gn_x)).sum()
    gJ, gh = torch.autograd.grad(total, (Jn, hn))
    return gJ.numpy(), gh.numpy()
