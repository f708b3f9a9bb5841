"""GPU: the reference's three LDS primitives (filter messages out, smoother / sampler from messages in)
through svae_lds_estep_fwd_ex, against the golden vectors of the reference's compiled Cython and the oracle.
"""
import numpy as np
import pytest
import torch
from oracle import lds_numpy as O
from util import golden_cases, load, rel_err, golden_natparams

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _gpu(x):
    return torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")


def _g(tree):
    return tuple(_gpu(x) for x in tree)


@pytest.mark.parametrize("path", golden_cases("lds_"))
def test_filter_messages_match_reference(path):
    """cython_lds_inference.pyx:83-86: prediction and filtered messages, natural form."""
    from svae_b200.lds import lds_inference as K
    d = load(path)
    init, pair, node = golden_natparams(d)
    ((Jp, hp), (Jf, hf)), lognorm = K.natural_filter_forward_general(_g(init), _g(pair), _g(node))
    for got, key in ((Jp, "J_pred"), (hp, "h_pred"), (Jf, "J_filt"), (hf, "h_filt"), (lognorm, "lognorm")):
        assert rel_err(got.cpu().numpy(), d[key]) < RTOL, key


@pytest.mark.parametrize("path", golden_cases("lds_"))
def test_smoother_and_sampler_from_reference_messages(path):
    """pyx:149-195 and 310-355 fed with the REFERENCE's messages reproduce the reference's statistics/samples."""
    from svae_b200.lds import lds_inference as K
    d = load(path)
    _, pair, _ = golden_natparams(d)
    msgs = ((_gpu(d["J_pred"]), _gpu(d["h_pred"])), (_gpu(d["J_filt"]), _gpu(d["h_filt"])))
    E_init, E_pair, E_node = K.natural_smoother_general(msgs, _g(pair))
    c = lambda x: x.cpu().numpy()
    assert rel_err(c(E_init[0]), d["E_init_xx"]) < RTOL and rel_err(c(E_init[1]), d["E_init_x"]) < RTOL
    for k, key in enumerate(("E_pair_xx", "E_pair_xxn", "E_pair_xnxn")):
        assert rel_err(c(E_pair[k]), d[key]) < RTOL, key
    assert rel_err(c(E_node[0]), d["E_node_xx"]) < RTOL and rel_err(c(E_node[1]), d["E_node_x"]) < RTOL
    S = d["eps"].shape[2]
    samples = K.natural_sample_backward(msgs, _g(pair), S, eps=_gpu(d["eps"]))
    assert rel_err(c(samples), d["samples"]) < RTOL


@pytest.mark.parametrize("n,T,B,S,inhomog", [(3, 6, 5, 2, False), (10, 40, 33, 1, False), (4, 9, 7, 3, True), (16, 12, 3, 1, False)])
def test_arbitrary_messages_vs_oracle(n, T, B, S, inhomog):
    """Messages that did NOT come from this filter (a different model's), with consistent predictions."""
    from svae_b200.lds import lds_inference as K
    rng = np.random.default_rng(100 + n)
    init, pair = O.rand_lds_natparam(n, rng)
    if inhomog:
        pairs = [O.rand_lds_natparam(n, rng)[1] for _ in range(T - 1)]
        pair = tuple(np.stack([p[k] for p in pairs]) for k in range(4))
    node = O.rand_node_potentials(B, T, n, rng)
    # filtered messages: random SPD precisions and random potentials; predictions derived from them
    R = rng.standard_normal((B, T, n, n))
    Jf = -0.5 * (R @ np.swapaxes(R, -1, -2) + n * np.eye(n))
    hf = rng.standard_normal((B, T, n))
    Jp, hp = np.empty_like(Jf), np.empty_like(hf)
    Jp[:, 0], hp[:, 0] = Jf[:, 0], hf[:, 0]                      # unused by both sides
    for t in range(T - 1):
        (Jp[:, t + 1], hp[:, t + 1]), _ = O.natural_predict(Jf[:, t], hf[:, t], *O._pair_at(pair, t))
    msgs = ((Jp, hp), (Jf, hf))
    eps = rng.standard_normal((B, T, S, n))
    want_stats = O.natural_smoother_general(msgs, pair)
    want_samples = O.natural_sample_backward_general(msgs, pair, eps)
    gm = tuple(_g(m) for m in msgs)
    E_init, E_pair, E_node = K.natural_smoother_general(gm, _g(pair))
    samples = K.natural_sample_backward(gm, _g(pair), S, eps=_gpu(eps))
    c = lambda x: x.cpu().numpy()
    for got, want in zip(E_init[:2] + E_pair[:3] + E_node[:2], want_stats[0][:2] + want_stats[1][:3] + want_stats[2][:2]):
        assert rel_err(c(got), want) < RTOL
    assert rel_err(c(samples), want_samples) < RTOL


def test_primitives_compose_to_the_fused_call():
    """lds_inference.py:253-264: filter -> sample_backward == natural_lds_sample; unbatched shapes."""
    from svae_b200.lds import lds_inference as K
    n, T, S = 5, 11, 2
    rng = np.random.default_rng(9)
    init, pair = O.rand_lds_natparam(n, rng)
    node = tuple(x[0] for x in O.rand_node_potentials(1, T, n, rng))
    eps = _gpu(rng.standard_normal((T, S, n)))
    msgs, lognorm = K.natural_filter_forward_general(_g(init), _g(pair), _g(node))
    assert msgs[0][0].shape == (T, n, n) and msgs[1][1].shape == (T, n) and lognorm.dim() == 0
    a = K.natural_sample_backward(msgs, _g(pair), S, eps=eps)
    b = K.natural_lds_sample((_g(init), _g(pair)), _g(node), S, eps=eps)
    assert a.shape == (T, S, n)
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < RTOL
    lz, stats = K.natural_lds_estep_general((_g(init), _g(pair)), _g(node))
    stats2 = K.natural_smoother_general(msgs, _g(pair))
    assert rel_err(lz.cpu().numpy(), lognorm.cpu().numpy()) < RTOL
    for x, y in zip(stats[1][:3], stats2[1][:3]):
        assert rel_err(x.cpu().numpy(), y.cpu().numpy()) < RTOL


def _dense_problem(n, T, B, S, inhomog, seed):
    rng = np.random.default_rng(seed)
    init, pair = O.rand_lds_natparam(n, rng)
    if inhomog:
        pairs = [O.rand_lds_natparam(n, rng)[1] for _ in range(T - 1)]
        pair = tuple(np.stack([p[k] for p in pairs]) for k in range(4))
    R = rng.standard_normal((B, T, n, n)) / np.sqrt(n)
    node_J = -0.5 * (R @ np.swapaxes(R, -1, -2) + 0.1 * np.eye(n))          # dense, natural form
    node = (node_J, rng.standard_normal((B, T, n)), 0.1 * rng.standard_normal((B, T)))
    eps = rng.standard_normal((B, T, S, n))
    return init, pair, node, eps, rng


@pytest.mark.parametrize("n,T,B,S,inhomog", [(2, 5, 3, 1, False), (10, 30, 37, 1, False), (5, 9, 6, 2, True), (16, 8, 3, 1, False)])
def test_dense_node_potentials_forward(n, T, B, S, inhomog):
    """lds_inference.py:65-80 ndims (3,2,1): dense (T,n,n) node precisions, full E[x x'] node statistic."""
    from svae_b200.lds import lds_inference as K
    init, pair, node, eps, _ = _dense_problem(n, T, B, S, inhomog, 300 + n)
    w_samples, (wE_init, wE_pair, wE_node), w_lognorm = O.natural_lds_inference_general((init, pair), node, eps)
    samples, (E_init, E_pair, E_node), lognorm = K.natural_lds_inference_general(
        (_g(init), _g(pair)), _g(node), num_samples=S, eps=_gpu(eps))
    c = lambda x: x.cpu().numpy()
    assert E_node[0].shape == (B, T, n, n)
    assert rel_err(c(lognorm), w_lognorm) < RTOL and rel_err(c(samples), w_samples) < RTOL
    assert rel_err(c(E_node[0]), wE_node[0]) < RTOL and rel_err(c(E_node[1]), wE_node[1]) < RTOL
    assert rel_err(c(E_init[0]), wE_init[0]) < RTOL and rel_err(c(E_init[1]), wE_init[1]) < RTOL
    for k in range(3):
        assert rel_err(c(E_pair[k]), wE_pair[k]) < RTOL
    # the messages too
    ((Jp, hp), (Jf, hf)), ln = K.natural_filter_forward_general(_g(init), _g(pair), _g(node))
    ((wJp, whp), (wJf, whf)), wln = O.natural_filter_forward_general(init, pair, node)
    assert rel_err(c(Jp), wJp) < RTOL and rel_err(c(hf), whf) < RTOL and rel_err(c(Jf), wJf) < RTOL and rel_err(c(ln), wln) < RTOL


@pytest.mark.parametrize("n,T,B,S,inhomog", [(3, 6, 4, 1, False), (10, 20, 9, 1, False), (4, 7, 5, 2, True)])
def test_dense_node_potentials_vjp(n, T, B, S, inhomog):
    """Gradient w.r.t. dense node precisions against torch autograd through the oracle (symmetric part:
    the port is symmetric, so only sym(dL/dJ) is defined)."""
    from svae_b200.lds import lds_inference as K
    from oracle import lds_torch as OT
    init, pair, node, eps, rng = _dense_problem(n, T, B, S, inhomog, 400 + n)
    t64 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)
    # oracle side
    Jn = t64(node[0]).requires_grad_(True); hn = t64(node[1]).requires_grad_(True)
    natparam = (tuple(t64(x) for x in init), tuple(t64(x) for x in pair))
    o_out = OT.natural_lds_inference_general(natparam, (Jn, hn, t64(node[2])), t64(eps))
    # ours
    gJn = _gpu(node[0]).requires_grad_(True); ghn = _gpu(node[1]).requires_grad_(True)
    k_out = K.natural_lds_inference_general((_g(init), _g(pair)), (gJn, ghn, _gpu(node[2])), num_samples=S, eps=_gpu(eps))

    def scalar(out, dev):
        samples, (E_init, E_pair, E_node), lognorm = out
        r = np.random.default_rng(5)
        tot = 0.
        for x in (samples, lognorm, E_init[0], E_init[1], E_pair[0], E_pair[1], E_pair[2], E_node[0], E_node[1]):
            w = torch.as_tensor(r.standard_normal(tuple(x.shape)), dtype=torch.float64, device=dev)
            tot = tot + (x * w).sum()
        return tot
    wJ, wh = torch.autograd.grad(scalar(o_out, "cpu"), (Jn, hn))
    kJ, kh = torch.autograd.grad(scalar(k_out, "cuda"), (gJn, ghn))
    sym = lambda x: 0.5 * (x + x.transpose(-1, -2))
    assert kJ.shape == (B, T, n, n)
    assert rel_err(kh.cpu().numpy(), wh.numpy()) < 1e-8
    assert rel_err(sym(kJ).cpu().numpy(), sym(wJ).numpy()) < 1e-8


# ------------------------------------------------------------------------------------------------
# the primitives' own VJPs (a4-a6): natural_filter_grad, natural_smoother_general_grad, natural_sample_backward_grad
# against the RAW outputs of the compiled reference (tests/golden/make_golden_prims.py)
# ------------------------------------------------------------------------------------------------
PRIM = golden_cases("ldsprim_")


def _leaf(x):
    return _gpu(x).requires_grad_(True)


@pytest.mark.parametrize("path", PRIM, ids=[p.split("/")[-1][:-4] for p in PRIM])
def test_filter_vjp_matches_reference_natural_filter_grad(path):
    """pyx:92-145: cotangents of BOTH message pairs and of lognorm in, cotangents of (J_node, h_node, logZ_node) out."""
    from svae_b200.lds import lds_inference as K
    d = load(path)
    init, pair, _ = golden_natparams(d)
    nJ, nh, nZ = _leaf(d["node_J"]), _leaf(d["node_h"]), _leaf(d["node_logZ"])
    ((Jp, hp), (Jf, hf)), ln = K.natural_filter_forward_general(_g(init), _g(pair), (nJ, nh, nZ))
    loss = (Jp * _gpu(d["in_gJp"])).sum() + (hp * _gpu(d["in_ghp"])).sum() + (Jf * _gpu(d["in_gJf"])).sum() \
        + (hf * _gpu(d["in_ghf"])).sum() + (ln * _gpu(d["in_gln"])).sum()
    loss.backward()
    assert rel_err(nJ.grad.cpu().numpy(), d["filt_gJ"]) < 1e-8
    assert rel_err(nh.grad.cpu().numpy(), d["filt_gh"]) < 1e-8
    assert rel_err(nZ.grad.cpu().numpy(), d["filt_glogZ"]) < 1e-12


@pytest.mark.parametrize("path", PRIM, ids=[p.split("/")[-1][:-4] for p in PRIM])
def test_smoother_and_sampler_vjps_compose_like_the_reference(path):
    """The reference returns cotangents for the prediction AND the filtered messages (its information-form RTS step reads
    both, pyx:236-306, 357-409); the covariance-form kernels read the filtered ones only.  For messages produced by the
    filter the two splits must agree once pushed through natural_filter_grad -- here: the reference's RAW message cotangents
    through this library's filter VJP (pinned above) against autograd through filter -> smoother + sampler."""
    from svae_b200.lds import lds_inference as K
    d = load(path)
    init, pair, _ = golden_natparams(d)
    S = d["eps"].shape[2]
    # (1) reference split, pushed through the filter VJP
    nJ, nh = _leaf(d["node_J"]), _leaf(d["node_h"])
    ((Jp, hp), (Jf, hf)), _ = K.natural_filter_forward_general(_g(init), _g(pair), (nJ, nh, _gpu(d["node_logZ"])))
    ((Jp * _gpu(d["sm_gJp"] + d["sa_gJp"])).sum() + (hp * _gpu(d["sm_ghp"] + d["sa_ghp"])).sum()
     + (Jf * _gpu(d["sm_gJf"] + d["sa_gJf"])).sum() + (hf * _gpu(d["sm_ghf"] + d["sa_ghf"])).sum()).backward()
    want_J, want_h = nJ.grad.clone(), nh.grad.clone()
    # (2) this library's primitives, differentiated one by one by autograd
    nJ2, nh2 = _leaf(d["node_J"]), _leaf(d["node_h"])
    msgs, _ = K.natural_filter_forward_general(_g(init), _g(pair), (nJ2, nh2, _gpu(d["node_logZ"])))
    E_init, E_pair, E_node = K.natural_smoother_general(msgs, _g(pair))
    samples = K.natural_sample_backward(msgs, _g(pair), S, eps=_gpu(d["eps"]))
    loss = (samples * _gpu(d["g_samples"])).sum() + (E_init[0] * _gpu(d["g_E_init_xx"])).sum() + (E_init[1] * _gpu(d["g_E_init_x"])).sum() \
        + (E_pair[0] * _gpu(d["g_E_pair_xx"])).sum() + (E_pair[1] * _gpu(d["g_E_pair_xxn"])).sum() \
        + (E_pair[2] * _gpu(d["g_E_pair_xnxn"])).sum() + (E_node[0] * _gpu(d["g_E_node_xx"])).sum() + (E_node[1] * _gpu(d["g_E_node_x"])).sum()
    loss.backward()
    assert rel_err(nJ2.grad.cpu().numpy(), want_J.cpu().numpy()) < 1e-8
    assert rel_err(nh2.grad.cpu().numpy(), want_h.cpu().numpy()) < 1e-8
    # the forward values on the way are the reference's
    assert rel_err(msgs[1][0].detach().cpu().numpy(), d["J_filt"]) < RTOL


def test_message_gradients_reach_a_custom_producer():
    """Messages that come from somewhere else (a leaf here): the smoother's VJP w.r.t. them against finite differences."""
    from svae_b200.lds import lds_inference as K
    d = load(PRIM[0])
    _, pair, _ = golden_natparams(d)
    Jf, hf = _leaf(d["J_filt"][:1]), _leaf(d["h_filt"][:1])
    w = _gpu(np.random.default_rng(0).standard_normal(d["h_filt"][:1].shape))
    f = lambda J, h: (K.natural_smoother_general(((None, None), (J, h)), _g(pair))[2][1] * w).sum()
    f(Jf, hf).backward()
    rng = np.random.default_rng(1)
    dh = _gpu(rng.standard_normal(hf.shape))
    dJ = _gpu(rng.standard_normal(Jf.shape)); dJ = dJ + dJ.transpose(-1, -2)
    eps = 1e-6
    with torch.no_grad():
        fd = (f(Jf + eps * dJ, hf + eps * dh) - f(Jf - eps * dJ, hf - eps * dh)) / (2 * eps)
    an = (Jf.grad * dJ).sum() + (hf.grad * dh).sum()
    assert abs(float(fd) - float(an)) < 1e-6 * max(1.0, abs(float(an)))


def test_primitives_above_the_compiled_latent_dimension():
    """n > 16: the three primitives run the same recursions as batched device linear algebra (svae_b200/lds/lds_large.py)
    behind the same names; filter differentiable w.r.t. the node potentials, smoother / sampler w.r.t. the messages."""
    from svae_b200.lds import lds_inference as K
    rng = np.random.default_rng(31)
    n, T, B, S = 19, 7, 3, 1
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    eps = rng.standard_normal((B, T, S, n))
    (w_pred, w_filt), w_lognorm = O.natural_filter_forward_general(init, pair, node)
    nJ, nh = _gpu(node[0]).requires_grad_(True), _gpu(node[1]).requires_grad_(True)
    ((Jp, hp), (Jf, hf)), lognorm = K.natural_filter_forward_general(_g(init), _g(pair), (nJ, nh))
    for got, want in ((Jp, w_pred[0]), (hp, w_pred[1]), (Jf, w_filt[0]), (hf, w_filt[1]), (lognorm, w_lognorm)):
        assert rel_err(got.detach().cpu().numpy(), want) < RTOL
    wE_init, wE_pair, wE_node = O.natural_smoother_general((w_pred, w_filt), pair)
    E_init, E_pair, E_node = K.natural_smoother_general(((Jp, hp), (Jf, hf)), _g(pair))
    assert rel_err(E_node[1].detach().cpu().numpy(), wE_node[1]) < RTOL and rel_err(E_pair[1].detach().cpu().numpy(), wE_pair[1]) < RTOL
    smp = K.natural_sample_backward(((Jp, hp), (Jf, hf)), _g(pair), S, _gpu(eps))
    assert rel_err(smp.detach().cpu().numpy(), O.natural_sample_backward_general((w_pred, w_filt), pair, eps)) < RTOL
    (smp.sum() + E_node[1].sum() - lognorm.sum()).backward()          # the chain filter -> smoother / sampler is differentiable
    assert bool(torch.isfinite(nJ.grad).all()) and bool(torch.isfinite(nh.grad).all()) and float(nh.grad.abs().max()) > 0
