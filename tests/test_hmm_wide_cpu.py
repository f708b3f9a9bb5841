"""CPU: the phase code of the wide HMM kernels (svae_b200/csrc/hmm_wide_core.h, K > 32 states) compiled for the host with
its threads looped (tests/emul/hmm_wide_host.cpp) reproduces the sequential oracle -- the arithmetic of svae_hmm_estep_wide
and of the K > 32 branch of svae_hmm_viterbi is pinned before any GPU runs it."""
import ctypes, os, subprocess
import numpy as np
import pytest
from oracle import hmm_numpy as HO
from util import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_host = None


def host():
    global _host
    if _host is None:
        out = os.path.join(ROOT, "tests", "emul", "_build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libhmm_wide_host.so")
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "svae_b200", "csrc"),
                               os.path.join(ROOT, "tests", "emul", "hmm_wide_host.cpp"), "-o", so])
        _host = ctypes.CDLL(so)
    return _host


def P_(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def wide_estep(lp, lA, node, fwd_only=False):
    node = np.ascontiguousarray(node, dtype=np.float64)
    B, T, K = node.shape
    lp, lA = np.ascontiguousarray(lp, dtype=np.float64), np.ascontiguousarray(lA, dtype=np.float64)
    logZ, E0, Et, Ez = np.zeros(B), np.full((B, K), np.nan), np.full((B, K, K), np.nan), np.full((B, T, K), np.nan)
    host().hmm_wide_estep_host(B, T, K, int(fwd_only), P_(lp), P_(lA), P_(node), P_(logZ), P_(E0), P_(Et), P_(Ez))
    return logZ, (E0, Et, Ez)


@pytest.mark.parametrize("K,T,B", [(33, 7, 2), (40, 50, 3), (64, 1, 2), (64, 2, 2), (65, 9, 1), (100, 23, 2), (129, 17, 1),
                                   (5, 12, 2), (32, 20, 2)])
def test_wide_estep_matches_oracle(K, T, B):
    rng = np.random.default_rng(K * 1000 + T)
    lp, lA = HO.rand_hmm_natparam(K, rng)
    lA = lA + rng.standard_normal((K, 1))                # unnormalised rows are allowed
    node = 2.0 * rng.standard_normal((B, T, K))
    w_logZ, (w0, wt, wz) = HO.hmm_estep((lp, lA, node))
    logZ, (E0, Et, Ez) = wide_estep(lp, lA, node)
    assert rel_err(logZ, w_logZ) < 1e-12
    assert rel_err(E0, w0) < 1e-11 and rel_err(Ez, wz) < 1e-11
    if T > 1:
        assert rel_err(Et, wt) < 1e-11
    else:
        assert np.all(Et == 0.0)
    assert np.allclose(Ez.sum(-1), 1.0, atol=1e-12)
    if T > 1:
        assert np.allclose(Et.sum((1, 2)), T - 1, rtol=1e-12)
    logZ2, _ = wide_estep(lp, lA, node, fwd_only=True)
    assert np.array_equal(logZ2, logZ)


def test_wide_estep_large_potentials_do_not_overflow():
    rng = np.random.default_rng(5)
    K, T, B = 48, 30, 2
    lp, lA = HO.rand_hmm_natparam(K, rng)
    node = 300.0 * rng.standard_normal((B, T, K))         # exp() of these overflows without the per-step shift
    w_logZ, (w0, wt, wz) = HO.hmm_estep((lp, lA, node))
    logZ, (E0, Et, Ez) = wide_estep(lp, lA, node)
    assert np.all(np.isfinite(logZ)) and rel_err(logZ, w_logZ) < 1e-12
    assert rel_err(Ez, wz) < 1e-10 and rel_err(Et, wt) < 1e-10


@pytest.mark.parametrize("K,T,B", [(33, 9, 2), (64, 40, 2), (100, 1, 1), (300, 12, 1)])
def test_wide_viterbi_matches_oracle(K, T, B):
    rng = np.random.default_rng(K + T)
    lp, lA = HO.rand_hmm_natparam(K, rng)
    node = 2.0 * rng.standard_normal((B, T, K))
    want = HO.hmm_viterbi((lp, lA, node))
    node = np.ascontiguousarray(node)
    path = np.zeros((B, T), dtype=np.int32)
    host().hmm_wide_viterbi_host(B, T, K, P_(np.ascontiguousarray(lp)), P_(np.ascontiguousarray(lA)), P_(node), P_(path))
    assert np.array_equal(path, want)
