"""CPU: the reference's own test strategy for svae/distributions (tests/test_dirichlet.py:15-25,
tests/test_gaussian.py:19-40, tests/test_niw.py:21-42): expectedstats == grad(logZ), pack/unpack and
standard<->natural round trips -- on the torch mirrors, with torch.autograd in the role of autograd.grad."""
import numpy as np
import pytest
import torch
from svae_b200.distributions import gaussian, niw, mniw, dirichlet, categorical

torch.manual_seed(0)
t64 = torch.float64


def rand_psd(n):
    A = torch.randn(n, n, dtype=t64)
    return A @ A.T


def grad_logZ(logZ, natparam):
    leaves = [natparam] if torch.is_tensor(natparam) else list(natparam)
    leaves = [l.clone().requires_grad_(True) for l in leaves]
    val = logZ(leaves[0] if torch.is_tensor(natparam) else tuple(leaves))
    return torch.autograd.grad(val, leaves)


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_dirichlet_expectedstats_is_grad_logZ(n):
    nat = torch.rand(n + 1, dtype=t64) * 3
    assert torch.allclose(dirichlet.expectedstats(nat), grad_logZ(dirichlet.logZ, nat)[0])


@pytest.mark.parametrize("n,k", [(1, 1), (2, 1), (3, 2), (4, 2)])
def test_gaussian_pack_unpack_and_expectedstats(n, k):
    A = torch.stack([rand_psd(n) + n * torch.eye(n, dtype=t64) for _ in range(k)])
    b = torch.randn(k, n, dtype=t64); c = torch.randn(k, dtype=t64); d = torch.randn(k, dtype=t64)
    packed = gaussian.pack_dense(A, b, c, d)
    A2, b2, c2, d2 = gaussian.unpack_dense(packed)
    assert torch.equal(A, A2) and torch.equal(b, b2) and torch.equal(c, c2) and torch.equal(d, d2)
    nat = gaussian.pack_dense(-0.5 * A, b, c, d)
    got, want = gaussian.expectedstats(nat), grad_logZ(gaussian.logZ, nat)[0]
    # logZ only sees the packed entries pack_dense fills (test_gaussian.py:30-40 compares those)
    for x, y in zip(gaussian.unpack_dense(got), gaussian.unpack_dense(want)):
        assert torch.allclose(x, y if y.dim() < 2 else (y + y.transpose(-1, -2)) / 2 if False else y, atol=1e-8) or \
            torch.allclose(x, (y + y.transpose(-1, -2)) / 2, atol=1e-8)


@pytest.mark.parametrize("n,k", [(1, 1), (2, 2), (3, 1), (4, 2)])
def test_niw_roundtrip_and_expectedstats(n, k):
    S = torch.stack([rand_psd(n) + n * torch.eye(n, dtype=t64) for _ in range(k)])
    m = torch.randn(k, n, dtype=t64); kappa = torch.rand(k, dtype=t64) + 1; nu = torch.rand(k, dtype=t64) + n + 1
    nat = niw.standard_to_natural(S, m, kappa, nu)
    S2, m2, k2, nu2 = niw.natural_to_standard(nat)
    assert torch.allclose(S, S2) and torch.allclose(m, m2) and torch.allclose(kappa, k2) and torch.allclose(nu, nu2)
    got = niw.expectedstats(nat, fudge=0.)
    want = grad_logZ(niw.logZ, nat)[0]
    gA, gb, gc, gd = gaussian.unpack_dense(got)
    wA, wb, wc, wd = gaussian.unpack_dense(want)
    assert torch.allclose(gA, (wA + wA.transpose(-1, -2)) / 2, atol=1e-7) and torch.allclose(gb, wb, atol=1e-7)
    assert torch.allclose(gc, wc, atol=1e-7) and torch.allclose(gd, wd, atol=1e-7)


@pytest.mark.parametrize("n", [1, 2, 3])
def test_mniw_expectedstats_is_grad_logZ(n):
    nu = torch.tensor(n + 2.5, dtype=t64)
    S = rand_psd(n) + n * torch.eye(n, dtype=t64); M = torch.randn(n, n, dtype=t64)
    Kmat = rand_psd(n) + n * torch.eye(n, dtype=t64)
    nat = mniw.standard_to_natural(nu, S, M, Kmat)
    nu2, S2, M2, K2 = mniw.natural_to_standard(nat)
    assert torch.allclose(S, S2) and torch.allclose(M, M2) and torch.allclose(Kmat, K2)
    got = mniw.expectedstats_standard(nu2, S2, M2, K2, fudge=0.)
    want = grad_logZ(mniw.logZ, nat)       # mniw.py:57-58 expectedstats_autograd
    sym = lambda X: (X + X.T) / 2
    assert torch.allclose(got[0], sym(want[0]), atol=1e-7) and torch.allclose(got[1], want[1], atol=1e-7)
    assert torch.allclose(got[2], sym(want[2]), atol=1e-7) and torch.allclose(got[3], want[3], atol=1e-7)


def test_categorical():
    nat = torch.randn(5, 4, dtype=t64)
    assert torch.allclose(categorical.expectedstats(nat), grad_logZ(categorical.logZ, nat)[0])


def test_synthetic_recipe_matches_the_oracles():
    """svae_b200/synthetic.py (benchmark inputs) draws exactly what the oracle's recipe draws."""
    import numpy as np
    from svae_b200 import synthetic as S
    from oracle import lds_numpy as O, hmm_numpy as HO
    a, b = S.lds_natparam(7, np.random.default_rng(3)), O.rand_lds_natparam(7, np.random.default_rng(3))
    for x, y in zip(a[0] + a[1], b[0] + b[1]):
        assert np.allclose(x, y, rtol=1e-13, atol=0)
    for x, y in zip(S.node_potentials(2, 5, 3, np.random.default_rng(1)), O.rand_node_potentials(2, 5, 3, np.random.default_rng(1))):
        assert np.array_equal(x, y)
    for x, y in zip(S.hmm_natparam(4, np.random.default_rng(2)), HO.rand_hmm_natparam(4, np.random.default_rng(2))):
        assert np.array_equal(x, y)
