"""GPU: parallel-in-time LDS E-step (svae_lds_compose_chunks + boundary chain + chunk launches) against the
sequential oracle -- pointwise, samples included (same noise => same paths)."""
import numpy as np
import pytest
import torch
from oracle import lds_numpy as O
from util import rel_err
import pit_numpy as P

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _gpu(x):
    return torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")


def _g(tree):
    return tuple(_gpu(x) for x in tree)


def _problem(n, T, B, S, seed):
    rng = np.random.default_rng(seed)
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    node = (node[0], node[1], 0.3 * rng.standard_normal((B, T)))
    return init, pair, node, rng.standard_normal((B, T, S, n))


@pytest.mark.parametrize("n,T,B,Lc", [(1, 9, 2, 3), (3, 14, 3, 4), (10, 41, 5, 8), (16, 20, 2, 7), (5, 12, 2, 1), (10, 203, 9, 16)])
def test_compose_kernel_matches_dense_algebra(n, T, B, Lc):
    from svae_b200.lds import lds_pit
    init, pair, node, _ = _problem(n, T, B, 1, 10 + n)
    want, w1, w2 = P.compose_numpy(pair, node, Lc)
    got, g1, g2 = lds_pit.compose_chunks(_g(pair), _g(node), Lc)
    for x, y in zip(want + (w1, w2), got + (g1, g2)):
        assert rel_err(y.cpu().numpy(), x) < 1e-10


@pytest.mark.parametrize("n,T,B,S,Lc", [
    (4, 25, 2, 2, 6),          # T-1 divisible by Lc
    (3, 14, 3, 1, 5),          # ragged tail chunk
    (10, 200, 4, 1, None),     # metric shape, default chunk length, probes ride in the statistics launch
    (16, 60, 2, 1, 9),         # n = 16: the unit probes need their own launches
    (8, 50, 3, 0, 7),          # statistics only
    (2, 400, 1, 3, 20),        # one long sequence
])
def test_parallel_equals_sequential(n, T, B, S, Lc):
    from svae_b200.lds import lds_pit
    init, pair, node, eps = _problem(n, T, B, S, 50 + n)
    w_samples, (wE_init, wE_pair, wE_node), w_lognorm = O.natural_lds_inference_general((init, pair), node, eps)
    samples, (E_init, E_pair, E_node), lognorm = lds_pit.natural_lds_inference_parallel(
        (_g(init), _g(pair)), _g(node), num_samples=S, eps=_gpu(eps) if S else None, chunk_len=Lc)
    c = lambda x: x.cpu().numpy()
    assert rel_err(c(lognorm), w_lognorm) < RTOL
    assert rel_err(c(E_init[0]), wE_init[0]) < RTOL and rel_err(c(E_init[1]), wE_init[1]) < RTOL
    for k in range(3):
        assert rel_err(c(E_pair[k]), wE_pair[k]) < RTOL, k
    assert np.allclose(c(E_pair[3]), T - 1.0)
    assert rel_err(c(E_node[0]), wE_node[0]) < RTOL and rel_err(c(E_node[1]), wE_node[1]) < RTOL
    if S:
        assert samples.shape == (B, T, S, n)
        assert rel_err(c(samples), w_samples) < 1e-8
    else:
        assert samples is None


def test_parallel_matches_the_fused_kernel_on_a_long_sequence():
    """T = 20000, B = 1: against this package's own sequential kernel (the oracle would take minutes)."""
    from svae_b200.lds import lds_pit, lds_inference as K
    n, T, B, S = 10, 20000, 1, 1
    init, pair, node, eps = _problem(n, T, B, S, 77)
    nat, nd, e = (_g(init), _g(pair)), _g(node), _gpu(eps)
    a = K.natural_lds_inference_general(nat, nd, S, e)
    b = lds_pit.natural_lds_inference_parallel(nat, nd, S, e)
    c = lambda x: x.cpu().numpy()
    assert rel_err(c(b[2]), c(a[2])) < RTOL and rel_err(c(b[0]), c(a[0])) < 1e-7
    for x, y in zip(a[1][1][:3] + a[1][2][:2], b[1][1][:3] + b[1][2][:2]):
        assert rel_err(c(y), c(x)) < RTOL


def test_unbatched_and_errors():
    from svae_b200.lds import lds_pit
    n, T = 3, 30
    init, pair, node, eps = _problem(n, T, 1, 1, 5)
    node1 = tuple(x[0] for x in node)
    w = O.natural_lds_inference_general((init, pair), node, eps)
    samples, stats, lognorm = lds_pit.natural_lds_inference_parallel((_g(init), _g(pair)), _g(node1), 1, _gpu(eps[0]), chunk_len=4)
    assert samples.shape == (T, 1, n) and lognorm.dim() == 0
    assert rel_err(samples.cpu().numpy(), w[0][0]) < 1e-8 and rel_err(lognorm.cpu().numpy(), w[2][0]) < RTOL
    with pytest.raises(ValueError):       # dense node potentials are not chunked
        lds_pit.natural_lds_inference_parallel((_g(init), _g(pair)), (_gpu(np.zeros((1, T, n, n))), _gpu(node[1])), 0)
    hn = _gpu(node[1]).requires_grad_(True)
    with pytest.raises(RuntimeError):     # forward-only
        lds_pit.natural_lds_inference_parallel((_g(init), _g(pair)), (_gpu(node[0]), hn), 0)


@pytest.mark.parametrize("per_seq", [False, True])
@pytest.mark.parametrize("n,T,B,S,Lc", [(3, 14, 2, 1, 5), (10, 61, 3, 1, 7), (6, 40, 2, 0, 6)])
def test_time_inhomogeneous_pair_parameters(per_seq, n, T, B, S, Lc):
    """Per-time and per-sequence pair parameters (what the SLDS feeds the LDS E-step, slds_svae.py:92-103)."""
    from svae_b200.lds import lds_pit
    rng = np.random.default_rng(30 + n)
    init, _ = O.rand_lds_natparam(n, rng)
    shape = (B, T - 1) if per_seq else (T - 1,)
    prs = [O.rand_lds_natparam(n, rng)[1] for _ in range(int(np.prod(shape)))]
    pair = tuple(np.stack([p[k] for p in prs]).reshape(shape + np.shape(prs[0][k])) for k in range(4))
    node = O.rand_node_potentials(B, T, n, rng)
    node = (node[0], node[1], 0.3 * rng.standard_normal((B, T)))
    eps = rng.standard_normal((B, T, S, n))
    want = [O.natural_lds_inference_general((init, tuple(x[q] for x in pair) if per_seq else pair),
                                            tuple(x[q:q + 1] for x in node), eps[q:q + 1]) for q in range(B)]
    cat = lambda f: np.concatenate([f(w) for w in want])
    samples, (E_init, E_pair, E_node), lognorm = lds_pit.natural_lds_inference_parallel(
        (_g(init), _g(pair)), _g(node), num_samples=S, eps=_gpu(eps) if S else None, chunk_len=Lc)
    c = lambda x: x.cpu().numpy()
    assert E_pair[1].shape == (B, T - 1, n, n)
    assert rel_err(c(lognorm), cat(lambda w: w[2])) < RTOL
    for k in range(3):
        assert rel_err(c(E_pair[k]), cat(lambda w: w[1][1][k])) < RTOL
    for k in range(2):
        assert rel_err(c(E_node[k]), cat(lambda w: w[1][2][k])) < RTOL
        assert rel_err(c(E_init[k]), cat(lambda w: w[1][0][k])) < RTOL
    if S:
        assert rel_err(c(samples), cat(lambda w: w[0])) < 1e-8


def test_non_positive_definite_is_reported():
    from svae_b200.lds import lds_pit
    n, T, B = 4, 40, 3
    init, pair, node, eps = _problem(n, T, B, 1, 9)
    bad = node[0].copy(); bad[1, 17] = +50.0              # a large negative precision in one sequence
    with pytest.raises(FloatingPointError):
        lds_pit.natural_lds_inference_parallel((_g(init), _g(pair)), (_gpu(bad), _gpu(node[1])), 0, chunk_len=6, check_status=True)
    lds_pit.natural_lds_inference_parallel((_g(init), _g(pair)), _g(node), 0, chunk_len=6, check_status=True)   # clean: no raise
