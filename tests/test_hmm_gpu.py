"""GPU: HMM forward-backward / logZ / Viterbi kernels against the golden vectors of the reference's compiled
cython_hmm_inference and against the NumPy oracle.  Moments to 1e-9 relative, Viterbi paths EXACT."""
import numpy as np
import pytest
import torch
from oracle import hmm_numpy as H
from util import golden_cases, load, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-9
_g = lambda x, grad=False: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda").requires_grad_(grad)


@pytest.mark.parametrize("path", golden_cases("hmm_"))
def test_hmm_golden(path):
    from svae_b200.hmm import hmm_inference as K
    d = load(path)
    nat = (_g(d["log_pi"]), _g(d["log_A"]), _g(d["node"]))
    logZ, (Ez0, Etrans, Ez) = K.hmm_estep(nat)
    assert rel_err(logZ.cpu().numpy(), d["logZ"]) < RTOL
    assert rel_err(Ez.cpu().numpy(), d["Ez"]) < RTOL and rel_err(Ez0.cpu().numpy(), d["Ez0"]) < RTOL
    if d["node"].shape[1] > 1:
        assert rel_err(Etrans.cpu().numpy(), d["Etrans"]) < RTOL
    assert np.array_equal(K.hmm_viterbi(nat).cpu().numpy(), d["viterbi"])
    assert rel_err(K.hmm_logZ(nat).cpu().numpy(), d["logZ"]) < RTOL


@pytest.mark.parametrize("K_,T,B", [(32, 500, 64), (8, 200, 37), (1, 4, 3), (2, 1, 5), (17, 33, 9), (5, 7, 1)])
def test_hmm_vs_oracle(K_, T, B):
    from svae_b200.hmm import hmm_inference as K
    rng = np.random.default_rng(K_ * 1000 + T)
    lp, lA = H.rand_hmm_natparam(K_, rng)
    node = 3.0 * rng.standard_normal((B, T, K_))
    w_logZ, (wE0, wEt, wEz) = H.hmm_estep((lp, lA, node))
    nat = (_g(lp), _g(lA), _g(node))
    logZ, (Ez0, Etrans, Ez) = K.hmm_estep(nat)
    assert rel_err(logZ.cpu().numpy(), w_logZ) < RTOL
    assert rel_err(Ez.cpu().numpy(), wEz) < RTOL and rel_err(Ez0.cpu().numpy(), wE0) < RTOL
    if T > 1:
        assert rel_err(Etrans.cpu().numpy(), wEt) < RTOL
    assert np.array_equal(K.hmm_viterbi(nat).cpu().numpy(), H.hmm_viterbi((lp, lA, node)))


@pytest.mark.parametrize("K_,T,B", [(33, 40, 5), (48, 1, 3), (64, 2, 2), (64, 300, 16), (100, 57, 4), (257, 21, 2)])
def test_hmm_wide_vs_oracle(K_, T, B):
    """More than 32 states: one CTA per sequence, thread = state (csrc/hmm_wide.cu; phase code pinned on CPU by
    tests/test_hmm_wide_cpu.py) -- E-step, log-normaliser alone, Viterbi, and the VJP of hmm_logZ."""
    from svae_b200.hmm import hmm_inference as K
    rng = np.random.default_rng(K_ * 1000 + T)
    lp, lA = H.rand_hmm_natparam(K_, rng)
    lA = lA + rng.standard_normal((K_, 1))
    node = 3.0 * rng.standard_normal((B, T, K_))
    w_logZ, (wE0, wEt, wEz) = H.hmm_estep((lp, lA, node))
    nat = (_g(lp), _g(lA), _g(node))
    logZ, (Ez0, Etrans, Ez) = K.hmm_estep(nat)
    assert rel_err(logZ.cpu().numpy(), w_logZ) < RTOL
    assert rel_err(Ez.cpu().numpy(), wEz) < RTOL and rel_err(Ez0.cpu().numpy(), wE0) < RTOL
    if T > 1:
        assert rel_err(Etrans.cpu().numpy(), wEt) < RTOL
    else:
        assert float(Etrans.abs().max()) == 0.0
    assert np.array_equal(K.hmm_viterbi(nat).cpu().numpy(), H.hmm_viterbi((lp, lA, node)))
    a, b_, c = _g(lp, True), _g(lA, True), _g(node, True)
    out = K.hmm_logZ((a, b_, c))
    assert rel_err(out.detach().cpu().numpy(), w_logZ) < RTOL
    out.sum().backward()
    assert rel_err(c.grad.cpu().numpy(), wEz) < RTOL and rel_err(a.grad.cpu().numpy(), wE0.sum(0)) < RTOL
    if T > 1:
        assert rel_err(b_.grad.cpu().numpy(), wEt.sum(0)) < RTOL


@pytest.mark.parametrize("K_,T,B", [(32, 500, 64), (32, 37, 13), (17, 33, 9), (24, 2, 8), (31, 1, 3), (20, 100, 1), (32, 9, 17)])
def test_hmm_tensor_core_path_equals_register_kernel(K_, T, B):
    """17..32 states: the tensor-core kernels (csrc/hmm_tc.cu, taken whenever the caller brings the workspace -- the Python
    wrapper always does) against the oracle AND against the register-resident kernel through its own C-ABI entry point;
    ragged batches (B not a multiple of the 8 sequences a warp owns), a per-sequence mask, forbidden transitions."""
    import ctypes
    from svae_b200 import _lib
    from svae_b200.hmm import hmm_inference as K
    rng = np.random.default_rng(K_ * 100 + T + B)
    lp, lA = H.rand_hmm_natparam(K_, rng)
    lA = lA + rng.standard_normal((K_, 1))
    lA[rng.random((K_, K_)) < 0.1] = -np.inf                   # forbidden transitions
    lA[np.arange(K_), np.arange(K_)] = 0.0
    node = 3.0 * rng.standard_normal((B, T, K_))
    w_logZ, (wE0, wEt, wEz) = H.hmm_estep((lp, lA, node))
    lpg, lAg, ng = _g(lp), _g(lA), _g(node)
    logZ, Ez0, Etrans, Ez = K.hmm_estep_forward(lpg, lAg, ng)
    assert rel_err(logZ.cpu().numpy(), w_logZ) < RTOL
    assert rel_err(Ez.cpu().numpy(), wEz) < RTOL and rel_err(Ez0.cpu().numpy(), wE0) < RTOL
    if T > 1:
        assert rel_err(Etrans.cpu().numpy(), wEt) < RTOL
    else:
        assert float(Etrans.abs().max()) == 0.0
    # the register-resident kernel on the same inputs
    new = lambda *sh: torch.empty(sh, dtype=torch.float64, device="cuda")
    l2, e02, et2, ez2 = new(B), new(B, K_), new(B, K_, K_), new(B, T, K_)
    P = _lib.ptr
    code = _lib.lib().svae_hmm_estep_masked(B, T, K_, 0, P(lpg), P(lAg), P(ng), None, P(l2), P(e02), P(et2), P(ez2),
                                            _lib.stream_ptr(ng.device))
    assert code == 0
    for a, b_ in ((logZ, l2), (Ez0, e02), (Etrans, et2), (Ez, ez2)):
        assert rel_err(a.cpu().numpy(), b_.cpu().numpy()) < 1e-12 or float(b_.abs().max()) == 0.0
    # log-normaliser alone
    lz, _, _, _ = K.hmm_estep_forward(lpg, lAg, ng, logz_only=True)
    assert rel_err(lz.cpu().numpy(), w_logZ) < RTOL
    # mask: skipped sequences keep what the output buffers held
    if B >= 3:
        active = torch.ones(B, dtype=torch.int32, device="cuda")
        active[1::2] = 0
        sent = tuple(torch.full_like(x, -7.0) for x in (logZ, Ez0, Etrans, Ez))
        got = K.hmm_estep_forward(lpg, lAg, ng, active=active, out=sent)
        for g_, f_ in zip(got, (logZ, Ez0, Etrans, Ez)):
            assert torch.equal(g_[0::2], f_[0::2]) and float((g_[1::2] + 7.0).abs().max()) == 0.0


def test_hmm_tensor_core_one_warp_sweep_in_a_child_process():
    """SVAE_HMM_TC_WARPS=1 selects the one-warp-per-8-sequences sweep (read once per process): same numbers."""
    import os, subprocess, sys
    from util import ROOT
    code = (
        "import numpy as np, torch\n"
        "from oracle import hmm_numpy as H\n"
        "from svae_b200.hmm import hmm_inference as K\n"
        "rng = np.random.default_rng(5)\n"
        "for K_, T, B in ((32, 40, 11), (19, 7, 8), (32, 1, 2)):\n"
        "    lp, lA = H.rand_hmm_natparam(K_, rng); node = 3.0 * rng.standard_normal((B, T, K_))\n"
        "    g = lambda x: torch.as_tensor(x, dtype=torch.float64, device='cuda')\n"
        "    logZ, (E0, Et, Ez) = K.hmm_estep((g(lp), g(lA), g(node)))\n"
        "    w_logZ, (w0, wt, wz) = H.hmm_estep((lp, lA, node))\n"
        "    err = lambda a, b: float(np.abs(a.cpu().numpy() - b).max() / max(np.abs(b).max(), 1e-300))\n"
        "    assert err(logZ, w_logZ) < 1e-9 and err(Ez, wz) < 1e-9 and err(E0, w0) < 1e-9 and (T == 1 or err(Et, wt) < 1e-9)\n"
        "print('ONE_WARP_OK')\n")
    env = dict(os.environ, SVAE_HMM_TC_WARPS="1", PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=180, cwd=ROOT)
    assert "ONE_WARP_OK" in r.stdout, r.stdout + r.stderr


def test_hmm_wide_mask_and_limits():
    from svae_b200.hmm import hmm_inference as K
    rng = np.random.default_rng(11)
    K_, T, B = 40, 9, 4
    lp, lA = H.rand_hmm_natparam(K_, rng)
    node = rng.standard_normal((B, T, K_))
    lpg, lAg, ng = _g(lp), _g(lA), _g(node)
    full = K.hmm_estep_forward(lpg, lAg, ng)
    sentinel = [torch.full_like(x, -7.0) for x in full]
    active = torch.tensor([1, 0, 1, 0], dtype=torch.int32, device="cuda")
    got = K.hmm_estep_forward(lpg, lAg, ng, active=active, out=tuple(sentinel))
    for g_, f_ in zip(got, full):
        assert torch.equal(g_[0], f_[0]) and torch.equal(g_[2], f_[2])
        assert float((g_[1] + 7.0).abs().max()) == 0.0 and float((g_[3] + 7.0).abs().max()) == 0.0
    with pytest.raises(ValueError):
        K.hmm_estep((torch.zeros(1025, device="cuda"), torch.zeros(1025, 1025, device="cuda"), torch.zeros(1, 2, 1025, device="cuda")))


def test_hmm_logZ_vjp_is_the_estep():
    from svae_b200.hmm import hmm_inference as K
    rng = np.random.default_rng(3)
    K_, T, B = 6, 12, 4
    lp, lA = H.rand_hmm_natparam(K_, rng)
    node = rng.standard_normal((B, T, K_))
    g = rng.standard_normal(B)
    a, b_, c = _g(lp, True), _g(lA, True), _g(node, True)
    (K.hmm_logZ((a, b_, c)) * _g(g)).sum().backward()
    _, (E0, Et, Ez) = H.hmm_estep((lp, lA, node))
    assert rel_err(a.grad.cpu().numpy(), (g[:, None] * E0).sum(0)) < RTOL
    assert rel_err(b_.grad.cpu().numpy(), (g[:, None, None] * Et).sum(0)) < RTOL
    assert rel_err(c.grad.cpu().numpy(), g[:, None, None] * Ez) < RTOL


def test_hmm_unbatched_like_reference():
    from svae_b200.hmm import hmm_inference as K
    rng = np.random.default_rng(4)
    lp, lA = H.rand_hmm_natparam(4, rng)
    node = rng.standard_normal((9, 4))
    logZ, (E0, Et, Ez) = K.hmm_estep((_g(lp), _g(lA), _g(node)))
    assert logZ.dim() == 0 and E0.shape == (4,) and Et.shape == (4, 4) and Ez.shape == (9, 4)
    assert K.hmm_viterbi((_g(lp), _g(lA), _g(node))).shape == (9,)


def test_hmm_forbidden_transitions():
    """Left-to-right chain: log A has -inf entries (zero-probability transitions) and log pi forbids all but state 0."""
    from svae_b200.hmm import hmm_inference as K
    rng = np.random.default_rng(11)
    K_, T, B = 6, 25, 4
    A = np.triu(rng.random((K_, K_)) + 0.1)
    A = A / A.sum(1, keepdims=True)
    with np.errstate(divide="ignore"):
        lA = np.log(A)
        lp = np.log(np.eye(K_)[0])
    node = rng.standard_normal((B, T, K_))
    w_logZ, (wE0, wEt, wEz) = H.hmm_estep((lp, lA, node))
    nat = (_g(lp), _g(lA), _g(node))
    logZ, (Ez0, Etrans, Ez) = K.hmm_estep(nat)
    assert rel_err(logZ.cpu().numpy(), w_logZ) < RTOL
    assert rel_err(Ez.cpu().numpy(), wEz) < RTOL and rel_err(Etrans.cpu().numpy(), wEt) < RTOL
    assert float(Etrans.cpu().numpy()[:, np.tril_indices(K_, -1)[0], np.tril_indices(K_, -1)[1]].max()) == 0.0
    path = K.hmm_viterbi(nat).cpu().numpy()
    assert np.array_equal(path, H.hmm_viterbi((lp, lA, node)))
    assert (np.diff(path, axis=1) >= 0).all() and (path[:, 0] == 0).all()
