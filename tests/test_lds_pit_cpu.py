"""CPU: the parallel-in-time pipeline (tests/pit_numpy.py) reproduces the sequential oracle pointwise, with the chunk
composition done (a) by dense NumPy algebra and (b) by the host build of the CUDA kernel's phase code."""
import numpy as np
import pytest
from oracle import lds_numpy as O
from util import rel_err
import pit_numpy as P


def _problem(n, T, B, S, seed):
    rng = np.random.default_rng(seed)
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    node = (node[0], node[1], 0.3 * rng.standard_normal((B, T)))
    return init, pair, node, rng.standard_normal((B, T, S, n))


@pytest.mark.parametrize("n,T,B,Lc", [(1, 9, 2, 3), (3, 14, 3, 4), (10, 41, 2, 8), (16, 20, 1, 7), (5, 12, 2, 1), (4, 10, 2, 50)])
def test_host_phase_code_matches_dense_algebra(n, T, B, Lc):
    init, pair, node, _ = _problem(n, T, B, 1, 10 + n)
    a, a1, a2 = P.compose_numpy(pair, node, Lc)
    b, b1, b2 = P.compose_host(pair, node, Lc)
    for x, y in zip(a + (a1, a2), b + (b1, b2)):
        assert rel_err(y, x) < 1e-11


@pytest.mark.parametrize("compose", [P.compose_numpy, P.compose_host])
@pytest.mark.parametrize("n,T,B,S,Lc", [(4, 25, 2, 2, 6), (3, 14, 2, 1, 5), (8, 30, 1, 1, 10)])
def test_pipeline_equals_sequential_oracle(compose, n, T, B, S, Lc):
    init, pair, node, eps = _problem(n, T, B, S, 20 + n)
    w_samples, (wE_init, wE_pair, wE_node), w_lognorm = O.natural_lds_inference_general((init, pair), node, eps)
    samples, (E_init, E_pair, E_node), lognorm = P.pit_estep(init, pair, node, eps, Lc, compose)
    tol = 1e-10
    assert rel_err(lognorm, w_lognorm) < tol and rel_err(samples, w_samples) < tol
    assert rel_err(E_init[0], wE_init[0]) < tol and rel_err(E_init[1], wE_init[1]) < tol
    for k in range(3):
        assert rel_err(E_pair[k], wE_pair[k]) < tol
    assert rel_err(E_node[0], wE_node[0]) < tol and rel_err(E_node[1], wE_node[1]) < tol


def _inhomog_problem(n, T, B, S, per_seq, seed):
    rng = np.random.default_rng(seed)
    init, _ = O.rand_lds_natparam(n, rng)
    shape = (B, T - 1) if per_seq else (T - 1,)
    pairs = [O.rand_lds_natparam(n, rng)[1] for _ in range(int(np.prod(shape)))]
    pair = tuple(np.stack([p[k] for p in pairs]).reshape(shape + np.shape(pairs[0][k])) for k in range(4))
    node = O.rand_node_potentials(B, T, n, rng)
    node = (node[0], node[1], 0.3 * rng.standard_normal((B, T)))
    return init, pair, node, rng.standard_normal((B, T, S, n))


@pytest.mark.parametrize("per_seq", [False, True])
@pytest.mark.parametrize("n,T,B,S,Lc", [(3, 14, 2, 1, 5), (6, 21, 2, 2, 4)])
def test_time_inhomogeneous_pair_parameters(per_seq, n, T, B, S, Lc):
    """Per-time and per-sequence pair parameters (the SLDS case, slds_svae.py:92-103) through the host phase code."""
    init, pair, node, eps = _inhomog_problem(n, T, B, S, per_seq, 30 + n)
    a, a1, a2 = P.compose_numpy(pair, node, Lc)
    b, b1, b2 = P.compose_host(pair, node, Lc)
    for x, y in zip(a + (a1, a2), b + (b1, b2)):
        assert rel_err(y, x) < 1e-11
    if per_seq:
        want = [O.natural_lds_inference_general((init, tuple(x[q] for x in pair)), tuple(x[q:q + 1] for x in node), eps[q:q + 1]) for q in range(B)]
        w_samples = np.concatenate([w[0] for w in want]); w_lognorm = np.concatenate([w[2] for w in want])
        wE_pair = tuple(np.concatenate([w[1][1][k] for w in want]) for k in range(3))
        wE_node = tuple(np.concatenate([w[1][2][k] for w in want]) for k in range(2))
    else:
        w_samples, (_, wE_pair, wE_node), w_lognorm = O.natural_lds_inference_general((init, pair), node, eps)
    samples, (E_init, E_pair, E_node), lognorm = P.pit_estep(init, pair, node, eps, Lc, P.compose_host)
    tol = 1e-10
    assert rel_err(lognorm, w_lognorm) < tol and rel_err(samples, w_samples) < tol
    for k in range(3):
        assert rel_err(E_pair[k], wE_pair[k]) < tol
    assert rel_err(E_node[0], wE_node[0]) < tol and rel_err(E_node[1], wE_node[1]) < tol
