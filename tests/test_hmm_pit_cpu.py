"""CPU: the parallel-in-time HMM pipeline reproduces the sequential oracle, with stages 1-2 done by log-space NumPy
and by the host build of the CUDA phase code (tests/emul/hmm_compose_host.cpp)."""
import numpy as np
import pytest
from oracle import hmm_numpy as HO
from util import rel_err
import hmm_pit_numpy as HP


@pytest.mark.parametrize("stage12", [HP.stage12_numpy, HP.stage12_host])
@pytest.mark.parametrize("K,T,B,Lc", [(3, 10, 2, 3), (8, 41, 3, 7), (32, 50, 2, 10), (5, 9, 2, 1), (4, 30, 1, 12), (32, 400, 1, 25)])
def test_pipeline_equals_sequential_oracle(stage12, K, T, B, Lc):
    rng = np.random.default_rng(K + T)
    lp, lA = HO.rand_hmm_natparam(K, rng)
    lA = lA + rng.standard_normal((K, 1))                # unnormalised rows are allowed
    node = 2.0 * rng.standard_normal((B, T, K))
    w_logZ, (w0, wt, wz) = HO.hmm_estep((lp, lA, node))
    logZ, (E0, Et, Ez) = HP.pit_estep(lp, lA, node, Lc, stage12)
    assert rel_err(logZ, w_logZ) < 1e-11
    assert rel_err(E0, w0) < 1e-10 and rel_err(Et, wt) < 1e-10 and rel_err(Ez, wz) < 1e-10


def test_messages_of_the_phase_code_match_log_space_algebra():
    rng = np.random.default_rng(3)
    K, T, B, Lc = 6, 23, 2, 5
    lp, lA = HO.rand_hmm_natparam(K, rng)
    node = rng.standard_normal((B, T, K))
    la, lb, lz = HP.stage12_numpy(lp, lA, node, Lc)
    ha, hb, hz = HP.stage12_host(lp, lA, node, Lc)
    assert np.allclose(ha, la, atol=1e-12) and np.allclose(hz, lz, atol=1e-11)
    assert np.allclose(hb - hb.max(-1, keepdims=True), lb - lb.max(-1, keepdims=True), atol=1e-11)


@pytest.mark.parametrize("stage12", [HP.viterbi_stage12_numpy, HP.viterbi_stage12_host])
@pytest.mark.parametrize("K,T,B,Lc", [(3, 10, 2, 3), (8, 41, 3, 7), (32, 50, 2, 10), (5, 9, 2, 1), (32, 400, 1, 25)])
def test_viterbi_pipeline_equals_sequential_oracle(stage12, K, T, B, Lc):
    """Max-plus composition + boundary back-tracking + pinned chunk decodes give the sequential Viterbi path."""
    rng = np.random.default_rng(100 + K + T)
    lp, lA = HO.rand_hmm_natparam(K, rng)
    node = 2.0 * rng.standard_normal((B, T, K))
    want = HO.hmm_viterbi((lp, lA, node))
    got = HP.pit_viterbi(lp, lA, node, Lc, stage12)
    assert np.array_equal(got, want)
