"""GMM local mean field (svae_b200/models/gmm.py, device-agnostic torch) against the NumPy restatement of
svae/models/gmm.py:62-128 on the pinwheel recipe of experiments/gmm_svae_synth.py:11-23."""
import numpy as np
import pytest
import torch
from oracle import gmm_numpy as G
from util import rel_err


def make_pinwheel_data(radial_std, tangential_std, num_classes, num_per_class, rate, rng):
    """experiments/gmm_svae_synth.py:11-23."""
    rads = np.linspace(0, 2 * np.pi, num_classes, endpoint=False)
    features = rng.standard_normal((num_classes * num_per_class, 2)) * np.array([radial_std, tangential_std])
    features[:, 0] += 1.
    labels = np.repeat(np.arange(num_classes), num_per_class)
    angles = rads[labels] + rate * np.exp(features[:, 0])
    rot = np.stack([np.cos(angles), -np.sin(angles), np.sin(angles), np.cos(angles)]).T.reshape((-1, 2, 2))
    return 10 * rng.permutation(np.einsum('ti,tij->tj', features, rot))


def _problem(K=10, N=2, T=50, seed=0):
    from svae_b200.models import gmm
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    pgm = gmm.init_pgm_param(K, N, alpha=0.05 / K, niw_conc=0.5, random_scale=3.0, generator=g)
    data = make_pinwheel_data(0.3, 0.05, 5, 100, 0.25, rng)[:T]
    node = (-0.5 * np.log1p(np.exp(rng.standard_normal((T, N)))), 0.3 * data[:, :N] + 0.1 * rng.standard_normal((T, N)))
    init = rng.random((T, K)); init /= init.sum(-1, keepdims=True)
    return pgm, node, init


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_local_meanfield_matches_oracle(device):
    from svae_b200.models import gmm
    pgm, node, init = _problem()
    want_local, want_prior, want_nat, want_kl, iters = G.local_meanfield(tuple(x.numpy() for x in pgm), node, init)
    assert 1 < iters < 100
    t = lambda x: torch.as_tensor(x, dtype=torch.float64, device=device)
    local, prior, nat, kl = gmm.local_meanfield(tuple(x.to(device) for x in pgm), tuple(t(x) for x in node), label_init=t(init))
    c = lambda x: x.detach().cpu().numpy()
    assert rel_err(c(local[0]), want_local[0]) < 1e-8 and rel_err(c(local[1]), want_local[1]) < 1e-8
    assert rel_err(c(prior[0]), want_prior[0]) < 1e-8 and rel_err(c(prior[1]), want_prior[1]) < 1e-8
    assert abs(float(kl) - want_kl) < 1e-8 * abs(want_kl)


def test_run_inference_is_differentiable_in_the_node_potentials_only():
    from svae_b200.models import gmm
    pgm, node, init = _problem(T=20)
    t = lambda x: torch.as_tensor(x, dtype=torch.float64)
    nJ, nh = t(node[0]).requires_grad_(True), t(node[1]).requires_grad_(True)
    samples, stats, global_kl, local_kl = gmm.run_inference(pgm, pgm, (nJ, nh), 3, label_init=t(init))
    assert samples.shape == (20, 3, 2) and abs(float(global_kl)) < 1e-9
    assert not stats[0].requires_grad and not stats[1].requires_grad          # unbox(stats), gmm.py:16
    (samples.sum() + local_kl).backward()
    assert nJ.grad is not None and torch.isfinite(nJ.grad).all() and torch.isfinite(nh.grad).all()


@pytest.mark.gpu
@pytest.mark.parametrize("K,N,T,B", [(10, 2, 50, 1), (15, 2, 50, 3), (4, 5, 33, 2), (7, 1, 300, 2)])
def test_cuda_fixed_point_kernel_matches_oracle(K, N, T, B):
    """csrc/gmm_meanfield.cu (whole coordinate ascent in one launch) against the NumPy restatement: same number of
    sweeps, same responsibilities."""
    from svae_b200.models import gmm
    from svae_b200.distributions import dirichlet, niw
    rng = np.random.default_rng(K * 10 + N)
    g = torch.Generator().manual_seed(K + N)
    pgm = gmm.init_pgm_param(K, N, alpha=0.05 / K, niw_conc=0.5, random_scale=3.0, generator=g)
    nJ = -0.5 * np.log1p(np.exp(rng.standard_normal((B, T, N))))
    nh = 2.0 * rng.standard_normal((B, T, N))
    init = rng.random((B, T, K)); init /= init.sum(-1, keepdims=True)
    t = lambda x: torch.as_tensor(x, dtype=torch.float64, device="cuda")
    lg = dirichlet.expectedstats(pgm[0].cuda()); gg = niw.expectedstats(pgm[1].cuda())
    labels, iters, kl = gmm.meanfield_fixed_point_cuda(lg, gg, t(nJ), t(nh), t(init))
    for b in range(B):
        want_local, _, _, _, w_iters = G.local_meanfield(tuple(x.numpy() for x in pgm), (nJ[b], nh[b]), init[b])
        assert int(iters[b]) == w_iters
        # the oracle re-evaluates the label factor once more at the optimum (gmm.py:76-77); compare that too
        local, _, _, _ = gmm.local_meanfield(tuple(x.cuda() for x in pgm), (t(nJ[b]), t(nh[b])), label_init=t(init[b]))
        assert rel_err(local[0].cpu().numpy(), want_local[0]) < 1e-8
        assert rel_err(local[1].cpu().numpy(), want_local[1]) < 1e-8
