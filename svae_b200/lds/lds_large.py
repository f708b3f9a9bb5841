"""LDS local E-step for latent dimensions ABOVE the compiled kernels' range (n > SVAE_LDS_MAX_N = 16).

The reference takes any n (LAPACK on n x n tiles, svae/lds/cython_lds_inference.pyx:28-90,149-210,310-355).  The fused sm_100a
kernels keep whole matrix columns in registers and are compiled for n = 1..16; beyond that the per-step matrices are large
enough for batched device linear algebra (B Cholesky factorisations / triangular solves / products of n x n per time step)
to be the sensible tool, so this module runs the SAME recursions as csrc/lds_estep_packed.cuh as batched torch operations on
the device, T steps in sequence, and lets autograd differentiate them (the reference differentiates w.r.t. the node
potentials only, lds_inference.py:27-29; so does this: the global parameters are detached).

    filter    M_t = J_pred + J11 + diag(J_node_t) = L L'      (information form: precision = -2 x natural parameter)
              G = M^-1 J12 (natural-form J12: the smoother gain),  m = M^-1 h_f
              J_pred' = J22 - J12' M^-1 J12,  h_pred' = J12' m,  lognorm += 1/2 h_f' m - sum log diag L
    smoother  C = G Sigma_{t+1};  Sigma_t = M^-1 + C G';  mu_t = m_t + G mu_{t+1}
    sampler   x_t = m_t + L^-T eps_t + G x_{t+1}               (eps[:, t] consumed at time t)

Public entry: estep(...) returns the dict lds_inference._pack_stats consumes.  Device tensors only, like every other
product path (no CPU path; the unit test drives the recursion on CPU tensors through the private _run).
"""
from __future__ import annotations
import torch
from .. import _lib


def _run(iJ, ih, ilogZ, J11, J12, J22, plogZ, nJ, nh, nlogZ, eps, S, per_t, want_stats=True, msgs_out=False, from_msgs=None):
    """All arguments canonical fp64 tensors on one device.  iJ (n,n)|(B,n,n), ih (n)|(B,n); pair blocks (n,n) | (T-1,n,n) |
    (B,T-1,n,n) natural form; nJ (B,T,n) diagonal or (B,T,n,n) dense natural form; nh (B,T,n); eps (B,T,S,n) or None.
    msgs_out: also return the prediction messages (natural form) as o["J_pred"], o["h_pred"] (natural_filter_forward_general).
    from_msgs = (J_filt (B,T,n,n), h_filt (B,T,n)) natural form: run the backward sweeps from GIVEN filtered messages
    (natural_smoother_general / natural_sample_backward; node potentials and init are then ignored, lognorm is meaningless)."""
    B, T, n = nh.shape
    dense = nJ.dim() == 4
    dev, dt = nh.device, nh.dtype
    homog = J11.dim() == 2

    def pair(t):              # natural-form blocks of step t, broadcastable against (B,n,n)
        if homog:
            return J11, J12, J22
        if J11.dim() == 3:
            return J11[t], J12[t], J22[t]
        return J11[:, t], J12[:, t], J22[:, t]

    a = (-2.0 * iJ).expand(B, n, n) if iJ.dim() == 2 else -2.0 * iJ
    hpred = ih.expand(B, n) if ih.dim() == 1 else ih
    if T > 1:
        a = a - 2.0 * pair(0)[0]
    lognorm = torch.zeros(B, dtype=dt, device=dev)
    info_max = torch.zeros((), dtype=torch.int32, device=dev)
    Gs, Minvs, ms, bs = [], [], [], []
    Jp, hp = [], []
    for t in range(T):
        if from_msgs is not None:
            M = -2.0 * from_msgs[0][:, t]
            if t < T - 1:
                M = M - 2.0 * pair(t)[0]
            hf = from_msgs[1][:, t]
        else:
            if msgs_out:      # a = J_pred (information form) + J11 of this step
                Jp.append(-0.5 * (a + 2.0 * pair(t)[0]) if t < T - 1 else -0.5 * a)
                hp.append(hpred)
            M = a + (-2.0 * nJ[:, t] if dense else torch.diag_embed(-2.0 * nJ[:, t]))
            hf = hpred + nh[:, t]
        L, info = torch.linalg.cholesky_ex(M)
        info_max = torch.maximum(info_max, info.max())
        eye = torch.eye(n, dtype=dt, device=dev).expand(B, n, n)
        Minv = torch.cholesky_solve(eye, L)
        m = torch.cholesky_solve(hf[..., None], L)[..., 0]
        lognorm = lognorm + 0.5 * (hf * m).sum(-1) - torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
        if S > 0:
            xi = torch.linalg.solve_triangular(L.transpose(-1, -2), eps[:, t].transpose(-1, -2), upper=True)   # (B,n,S)
            bs.append(m[..., None] + xi)
        ms.append(m)
        Minvs.append(Minv)
        if t < T - 1:
            p11, p12, p22 = pair(t)
            G = torch.cholesky_solve(p12.expand(B, n, n), L)             # M^-1 J12 (natural-form J12): the smoother gain
            Gs.append(G)
            nxt = -2.0 * p22
            if t + 1 < T - 1:
                nxt = nxt - 2.0 * pair(t + 1)[0]
            a = nxt - p12.transpose(-1, -2) @ G
            hpred = (p12.transpose(-1, -2) @ m[..., None])[..., 0]

    # backward in time
    Sig, mu = Minvs[T - 1], ms[T - 1]
    x = bs[T - 1] if S > 0 else None
    Exx = [None] * T
    Exxn = [None] * (T - 1)
    mus = [None] * T
    xs = [None] * T
    mus[T - 1], xs[T - 1] = mu, x
    Exx[T - 1] = Sig + mu[..., :, None] * mu[..., None, :]
    for t in range(T - 2, -1, -1):
        G = Gs[t]
        C = G @ Sig
        mun = mu
        mu = ms[t] + (G @ mun[..., None])[..., 0]
        if S > 0:
            x = bs[t] + G @ x
        Sig = Minvs[t] + C @ G.transpose(-1, -2)
        Exx[t] = Sig + mu[..., :, None] * mu[..., None, :]
        Exxn[t] = C + mu[..., :, None] * mun[..., None, :]
        mus[t], xs[t] = mu, x

    konst = torch.zeros(B, dtype=dt, device=dev)
    if ilogZ is not None:
        konst = konst + ilogZ.reshape(-1)
    if plogZ is not None and T > 1:
        if homog:
            konst = konst + (T - 1.0) * plogZ.reshape(())
        elif plogZ.dim() == 1:
            konst = konst + plogZ.sum()
        else:
            konst = konst + plogZ.sum(-1)
    if nlogZ is not None:
        konst = konst + nlogZ.sum(-1)
    Exx_t = torch.stack(Exx, 1)                                           # (B,T,n,n)
    E_x = torch.stack(mus, 1)
    o = dict(B=B, T=T, n=n, S=S, eps=eps, dense=dense, lognorm=lognorm + konst, status=info_max)
    o["E_node_x"] = E_x
    o["E_node_xx"] = Exx_t if dense else torch.diagonal(Exx_t, dim1=-2, dim2=-1)
    o["samples"] = torch.stack(xs, 1).transpose(-1, -2) if S > 0 else None        # (B,T,S,n)
    if msgs_out and from_msgs is None:
        o["J_pred"], o["h_pred"] = torch.stack([x.expand(B, n, n) for x in Jp], 1), torch.stack([x.expand(B, n) for x in hp], 1)
    if want_stats:
        o["E_init_xx"], o["E_init_x"] = Exx[0], mus[0]
        if T > 1:
            Exxn_t = torch.stack(Exxn, 1)
            o["E_pair_xx"], o["E_pair_xxn"], o["E_pair_xnxn"] = Exx_t[:, :-1].sum(1), Exxn_t.sum(1), Exx_t[:, 1:].sum(1)
        else:
            Exxn_t = torch.zeros((B, 0, n, n), dtype=dt, device=dev)
            z = torch.zeros((B, n, n), dtype=dt, device=dev)
            o["E_pair_xx"], o["E_pair_xxn"], o["E_pair_xnxn"] = z, z.clone(), z.clone()
        if per_t or dense:
            o["Exx_t"] = Exx_t
        if per_t:
            o["Exxn_t"] = Exxn_t
    return o


def estep(init_params, pair_params, node_params, num_samples, eps, per_t, generator=None, want_stats=True, check_status=False):
    """Same contract as lds_inference._estep (a dict of batched tensors; gradients reach the node potentials)."""
    from . import lds_inference as LI
    h_any = node_params[1]
    device = h_any.device if torch.is_tensor(h_any) and h_any.is_cuda else torch.device("cuda", torch.cuda.current_device())
    nJ, nh, nlogZ, batched = LI._canonical_node_params(node_params, device)
    if not nh.is_cuda:
        raise RuntimeError("svae_b200: the LDS E-step runs on CUDA tensors only (there is no CPU path)")
    B, T, n = nh.shape
    iJ, ih, ilogZ = LI._canonical_init_params(init_params, device)
    if iJ.dim() == 3 and (iJ.shape != (B, n, n) or ih.shape != (B, n)):
        raise ValueError("per-sequence init_params shapes")
    if iJ.dim() == 2 and (iJ.shape != (n, n) or ih.shape != (n,)):
        raise ValueError("init_params shapes")
    J11, J12, J22, plogZ, _, _ = LI._canonical_pair_params(pair_params, B, T, n, device)
    S = int(num_samples) if num_samples else 0
    if eps is not None:
        eps = _lib.f64(eps, device)
        if not batched and eps.dim() == 3:
            eps = eps[None]
        if eps.shape != (B, T, S, n):
            raise ValueError("eps must be (B,T,S,n) = %s, got %s" % ((B, T, S, n), tuple(eps.shape)))
    elif S > 0:
        eps = torch.randn((B, T, S, n), dtype=torch.float64, device=device, generator=generator)
    det = lambda v: None if v is None else v.detach()
    o = _run(det(iJ), det(ih), det(ilogZ), det(J11), det(J12), det(J22), det(plogZ), nJ, nh, nlogZ, det(eps), S, per_t, want_stats)
    o["batched"] = batched
    if check_status and int(o["status"].item()):
        raise FloatingPointError("LDS E-step: a filtered precision is not positive definite (reference asserts: "
                                 "svae/lds/gaussian.py:48,64)")
    return o


def filter_forward(init_params, pair_params, node_params):
    """natural_filter_forward_general for n > 16: ((J_pred, h_pred), (J_filt, h_filt)) natural form, lognorm; batched."""
    from . import lds_inference as LI
    h_any = node_params[1]
    device = h_any.device if torch.is_tensor(h_any) and h_any.is_cuda else torch.device("cuda", torch.cuda.current_device())
    nJ, nh, nlogZ, batched = LI._canonical_node_params(node_params, device)
    if not nh.is_cuda:
        raise RuntimeError("svae_b200: the LDS E-step runs on CUDA tensors only (there is no CPU path)")
    B, T, n = nh.shape
    iJ, ih, ilogZ = LI._canonical_init_params(init_params, device)
    J11, J12, J22, plogZ, _, _ = LI._canonical_pair_params(pair_params, B, T, n, device)
    det = lambda v: None if v is None else v.detach()
    o = _run(det(iJ), det(ih), det(ilogZ), det(J11), det(J12), det(J22), det(plogZ), nJ, nh, nlogZ, None, 0, False,
             want_stats=False, msgs_out=True)
    J_filt = o["J_pred"] + (nJ if nJ.dim() == 4 else torch.diag_embed(nJ))
    return ((o["J_pred"], o["h_pred"]), (J_filt, o["h_pred"] + nh)), o["lognorm"], batched


def from_messages(J_filt, h_filt, pair_params, num_samples, eps, per_t, want_stats, generator=None):
    """natural_smoother_general / natural_sample_backward for n > 16 from given filtered messages (B,T,n,n), (B,T,n):
    the dict lds_inference._pack_stats consumes; gradients reach the messages through autograd."""
    from . import lds_inference as LI
    if not h_filt.is_cuda:
        raise RuntimeError("svae_b200: the LDS E-step runs on CUDA tensors only (there is no CPU path)")
    B, T, n = h_filt.shape
    device = h_filt.device
    J11, J12, J22, plogZ, _, _ = LI._canonical_pair_params(pair_params, B, T, n, device)
    S = int(num_samples) if num_samples else 0
    if J_filt.shape != (B, T, n, n):
        raise ValueError("forward messages must be ((B,T,n,n), (B,T,n))")
    if eps is not None:
        eps = _lib.f64(eps, device)
        if eps.shape != (B, T, S, n):
            raise ValueError("eps must be (B,T,S,n) = %s, got %s" % ((B, T, S, n), tuple(eps.shape)))
    elif S > 0:
        eps = torch.randn((B, T, S, n), dtype=torch.float64, device=device, generator=generator)
    det = lambda v: None if v is None else v.detach()
    z = torch.zeros((B, T, n), dtype=torch.float64, device=device)
    eye = -0.5 * torch.eye(n, dtype=torch.float64, device=device)
    o = _run(eye, z[0, 0], None, det(J11), det(J12), det(J22), None, z, z, None, det(eps), S, per_t, want_stats,
             from_msgs=(J_filt, h_filt))
    o["lognorm"] = torch.zeros(B, dtype=torch.float64, device=device)
    return o
