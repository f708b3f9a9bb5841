"""CPU: pins oracle/hmm_numpy.py to the reference's compiled hmm_logZ / hmm_logZ_grad (golden + live)."""
import numpy as np
import pytest
from oracle import hmm_numpy as H, ref
from util import golden_cases, load, rel_err


@pytest.mark.parametrize("path", golden_cases("hmm_"))
def test_hmm_oracle_matches_golden(path):
    d = load(path)
    logZ, (Ez0, Etrans, Ez) = H.hmm_estep((d["log_pi"], d["log_A"], d["node"]))
    assert rel_err(logZ, d["logZ"]) < 1e-12 and rel_err(Ez, d["Ez"]) < 1e-10
    assert rel_err(Ez0, d["Ez0"]) < 1e-10
    if d["node"].shape[1] > 1:
        assert rel_err(Etrans, d["Etrans"]) < 1e-10
    assert np.array_equal(H.hmm_viterbi((d["log_pi"], d["log_A"], d["node"])), d["viterbi"])


def test_hmm_stats_sum_rules():
    rng = np.random.default_rng(1)
    K, T, B = 6, 11, 4
    lp, lA = H.rand_hmm_natparam(K, rng)
    node = rng.standard_normal((B, T, K))
    logZ, (Ez0, Etrans, Ez) = H.hmm_estep((lp, lA, node))
    assert np.allclose(Ez.sum(-1), 1.0) and np.allclose(Etrans.sum((1, 2)), T - 1)
    assert np.allclose(Etrans.sum(2), Ez[:, :-1].sum(1)) and np.allclose(Etrans.sum(1), Ez[:, 1:].sum(1))


def test_viterbi_is_the_best_path_bruteforce():
    import itertools
    rng = np.random.default_rng(2)
    K, T = 3, 5
    lp, lA = H.rand_hmm_natparam(K, rng)
    node = rng.standard_normal((1, T, K))
    best, arg = -np.inf, None
    for path in itertools.product(range(K), repeat=T):
        sc = lp[path[0]] + node[0, 0, path[0]] + sum(lA[path[t - 1], path[t]] + node[0, t, path[t]] for t in range(1, T))
        if sc > best:
            best, arg = sc, path
    assert tuple(H.hmm_viterbi((lp, lA, node))[0]) == arg
