"""Recognition / decoder networks -- torch mirror of the reference's svae/nnet.py (autograd MLP and "gresnet").

Out of scope as a kernel target: these are dense GEMMs and stay ordinary PyTorch (cuBLAS
on the tensor cores); only their OUTPUT CONTRACT matters to the E-step: `gaussian_info` emits node potentials
(J = -1/2 softplus(.), h) (nnet.py:43-47).  Parameters are nested tuples of tensors like the reference's.
"""
from __future__ import annotations
import math
import torch
from .util import log1pexp, sigmoid, relu, identity  # noqa: F401  (re-exported like the reference)

tanh = torch.tanh


def rand_partial_isometry(m, n, generator=None, device=None):
    """nnet.py:13-15."""
    d = max(m, n)
    q = torch.linalg.qr(torch.randn(d, d, dtype=torch.float64, generator=generator))[0][:m, :n]
    return q.to(device) if device is not None else q


def _ravelers(shape):
    ravel = lambda x: x.reshape(-1, shape[-1])
    unravel = lambda y: y.reshape(tuple(shape[:-1]) + (-1,))
    return ravel, unravel


def layer(nonlin, W, b, inputs):
    """nnet.py:25."""
    return nonlin(inputs @ W + b)


def init_layer(d_in, d_out, scale=1e-2, generator=None, device=None):
    """nnet.py:26-30 (init_layer_random)."""
    W = scale * torch.randn(d_in, d_out, dtype=torch.float64, generator=generator)
    b = scale * torch.randn(d_out, dtype=torch.float64, generator=generator)
    return (W.to(device), b.to(device)) if device is not None else (W, b)


def gaussian_mean(inputs, sigmoid_mean=False):
    """nnet.py:35-41."""
    mu_input, sigmasq_input = torch.chunk(inputs, 2, dim=-1)
    mu = torch.sigmoid(mu_input) if sigmoid_mean else mu_input
    return (mu, log1pexp(sigmasq_input))


def gaussian_info(inputs):
    """nnet.py:43-47: the node-potential head (J = -1/2 softplus, h)."""
    J_input, h = torch.chunk(inputs, 2, dim=-1)
    return (-0.5 * log1pexp(J_input), h)


def _mlp(nonlinearities):
    def mlp(params, inputs):
        ravel, unravel = _ravelers(inputs.shape)
        out = ravel(inputs)
        for nonlin, (W, b) in zip(nonlinearities, params):
            out = layer(nonlin, W, b, out)
        return unravel(out) if torch.is_tensor(out) else tuple(unravel(o) for o in out)

    def pre_head(params, inputs):
        """activations entering the output head: every layer but the last nonlinearity, raveled (rows, d_out)."""
        out = _ravelers(inputs.shape)[0](inputs)
        for nonlin, (W, b) in zip(nonlinearities[:-1], params[:-1]):
            out = layer(nonlin, W, b, out)
        W, b = params[-1]
        return out @ W + b
    mlp.pre_head = pre_head
    return mlp


class _InfoHeadFn(torch.autograd.Function):
    """gaussian_info + the resnet's 'info' residual in ONE kernel (csrc/nnet_heads.cu): the node potentials of the E-step
    are written directly; backward is one kernel plus the two small reductions for the residual parameters."""

    @staticmethod
    def forward(ctx, act, res, b2):
        from . import _lib
        rows, two_n = act.shape
        n = two_n // 2
        act, res, b2 = act.contiguous(), res.contiguous(), b2.contiguous()
        J = torch.empty((rows, n), dtype=torch.float64, device=act.device)
        h = torch.empty((rows, n), dtype=torch.float64, device=act.device)
        P = _lib.ptr
        _lib.check(_lib.lib().svae_gaussian_info_head_fwd(rows, n, P(act), P(res), P(b2), P(J), P(h), _lib.stream_ptr(act.device)),
                   "svae_gaussian_info_head_fwd")
        ctx.save_for_backward(act, b2)
        return J, h

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gJ, gh):
        from . import _lib
        act, b2 = ctx.saved_tensors
        rows, two_n = act.shape
        n = two_n // 2
        gJ = None if gJ is None else gJ.contiguous()
        gh = None if gh is None else gh.contiguous()
        g_act = torch.empty_like(act)
        P = _lib.ptr
        _lib.check(_lib.lib().svae_gaussian_info_head_bwd(rows, n, P(act), P(gJ), P(gh), P(g_act), _lib.stream_ptr(act.device)),
                   "svae_gaussian_info_head_bwd")
        g_b2 = None if gJ is None else -0.5 * torch.sigmoid(b2) * gJ.sum(0)
        return g_act, gh, g_b2


def init_mlp(d_in, layer_specs, generator=None, device=None):
    """nnet.py:60-65: layer_specs = [(d_out, nonlinearity), ...] -> (mlp(params, inputs), params)."""
    dims = [d_in] + [l[0] for l in layer_specs]
    params = tuple(init_layer(a, b, generator=generator, device=device) for a, b in zip(dims[:-1], dims[1:]))
    return _mlp([l[1] for l in layer_specs]), params


def _diagonal_gaussian_loglike(x, mu, sigmasq):
    """nnet.py:70-75, with an optional leading batch axis: x (...,T,p), mu (...,T,K,p) or (...,T,p)."""
    if mu.dim() == x.dim():
        mu, sigmasq = mu.unsqueeze(-2), sigmasq.unsqueeze(-2)
    K = mu.shape[-2]
    Tp = x.numel()
    return -Tp / 2. * math.log(2 * math.pi) + (-0.5 * ((x.unsqueeze(-2) - mu) ** 2 / sigmasq).sum()
                                              - 0.5 * torch.log(sigmasq).sum()) / K


def make_loglike(gaussian_mlp):
    """nnet.py:77-80."""
    def loglike(params, inputs, targets):
        return _diagonal_gaussian_loglike(targets, *gaussian_mlp(params, inputs))
    return loglike


def _gresnet(mlp_type, mlp):
    """nnet.py:90-104: MLP plus a linear residual path."""
    def gresnet(params, inputs):
        ravel, unravel = _ravelers(inputs.shape)
        mlp_params, (W, b1, b2) = params
        if mlp_type == 'mean':
            mu_mlp, sigmasq_mlp = mlp(mlp_params, inputs)
            mu_res = unravel(ravel(inputs) @ W + b1)
            return (mu_mlp + mu_res, sigmasq_mlp + log1pexp(b2))
        if inputs.is_cuda and inputs.dtype == torch.float64 and hasattr(mlp, "pre_head"):
            # fused output head: one kernel from the last layer's activations to (J_node, h_node)
            J, h = _InfoHeadFn.apply(mlp.pre_head(mlp_params, inputs), ravel(inputs) @ W + b1, b2)
            return (unravel(J), unravel(h))
        J_mlp, h_mlp = mlp(mlp_params, inputs)
        h_res = unravel(ravel(inputs) @ W + b1)
        return (J_mlp + (-0.5 * log1pexp(b2)), h_mlp + h_res)
    return gresnet


def init_gresnet(d_in, layer_specs, generator=None, device=None):
    """nnet.py:106-110."""
    d_out = layer_specs[-1][0] // 2
    mlp_type = 'info' if layer_specs[-1][1] is gaussian_info else 'mean'
    res = (rand_partial_isometry(d_in, d_out, generator, device),
           torch.zeros(d_out, dtype=torch.float64, device=device), torch.zeros(d_out, dtype=torch.float64, device=device))
    mlp, mlp_params = init_mlp(d_in, layer_specs, generator=generator, device=device)
    return _gresnet(mlp_type, mlp), (mlp_params, res)
