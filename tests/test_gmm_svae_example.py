"""f1: the GMM-SVAE experiment of the reference (experiments/gmm_svae_synth.py:25-61) as examples/gmm_svae_synth.py -- a short
run must make the objective fall; on the GPU the step uses svae_gmm_meanfield inside make_gradfun and the fused
gaussian_info head (csrc/nnet_heads.cu), checked against the plain torch head."""
import importlib.util, os
import numpy as np
import pytest
import torch
from util import ROOT


def _example():
    spec = importlib.util.spec_from_file_location("gmm_svae_synth", os.path.join(ROOT, "examples", "gmm_svae_synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _check(log):
    assert len(log) == 60 and all(np.isfinite(v) for v in log)
    assert sum(log[-10:]) / 10 < 0.5 * sum(log[:10]) / 10, (log[:3], log[-3:])


def test_gmm_svae_short_run_converges_cpu():
    log, _ = _example().main(iters=60, device="cpu", quiet=True)
    _check(log)


@pytest.mark.gpu
def test_gmm_svae_short_run_converges_gpu():
    from svae_b200 import _lib
    n0 = _lib.launch_count()
    log, params = _example().main(iters=60, device="cuda", quiet=True)
    _check(log)
    assert _lib.launch_count() - n0 >= 3 * 60        # fused head fwd + bwd and the mean-field kernel, every iteration
    assert all(torch.isfinite(p).all() for p in params[0])


@pytest.mark.gpu
def test_fused_info_head_matches_the_torch_head():
    """svae/nnet.py:43-47,90-104: values and gradients of (J_node, h_node) w.r.t. every network parameter."""
    from svae_b200 import nnet
    g = torch.Generator().manual_seed(0)
    net, params = nnet.init_gresnet(6, [(16, torch.tanh), (2 * 5, nnet.gaussian_info)], generator=g, device="cuda")
    x = torch.randn(7, 11, 6, dtype=torch.float64, generator=g).cuda()
    wJ, wh = torch.randn(7, 11, 5, dtype=torch.float64, generator=g).cuda(), torch.randn(7, 11, 5, dtype=torch.float64, generator=g).cuda()
    leaves = lambda p: [t.detach().clone().requires_grad_(True) for t in nnet_leaves(p)]

    def nnet_leaves(p):
        (mlp_params, res) = p
        return [t for Wb in mlp_params for t in Wb] + list(res)

    def rebuild(ls):
        k = len(params[0])
        return (tuple((ls[2 * i], ls[2 * i + 1]) for i in range(k)), tuple(ls[2 * k:]))
    la = leaves(params)
    J, h = net(rebuild(la), x)                       # CUDA fp64: the fused head
    ((J * wJ).sum() + (h * wh).sum()).backward()
    lb = [t.detach().cpu().clone().requires_grad_(True) for t in la]
    J2, h2 = net(rebuild(lb), x.cpu())               # CPU: the torch head
    ((J2 * wJ.cpu()).sum() + (h2 * wh.cpu()).sum()).backward()
    assert torch.allclose(J.cpu(), J2, rtol=1e-12, atol=1e-14) and torch.allclose(h.cpu(), h2, rtol=1e-12, atol=1e-14)
    for a, b in zip(la, lb):
        assert torch.allclose(a.grad.cpu(), b.grad, rtol=1e-10, atol=1e-12)
