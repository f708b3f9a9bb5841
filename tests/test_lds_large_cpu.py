"""CPU: the recursion behind the large-n LDS E-step (svae_b200/lds/lds_large.py, latent dimensions above the compiled kernels'
16) against the oracle -- forward values and, through autograd, the VJP w.r.t. the node potentials.  The public entry
point takes device tensors only; this test drives the private recursion on CPU tensors."""
import numpy as np
import pytest
import torch
from oracle import lds_numpy as O, lds_torch as OT
from util import rel_err
from svae_b200.lds import lds_large as LL

t64 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)


def run(init, pair, node, eps, S, per_t=False, grad=False):
    iJ, ih, ilz = (t64(x) for x in init)
    J11, J12, J22, plz = (t64(x) for x in pair)
    nJ, nh = t64(node[0]).requires_grad_(grad), t64(node[1]).requires_grad_(grad)
    nlz = t64(node[2]) if len(node) > 2 else None
    o = LL._run(iJ, ih, ilz.reshape(-1), J11, J12, J22, plz, nJ, nh, nlz, None if eps is None else t64(eps), S, per_t)
    return o, nJ, nh


@pytest.mark.parametrize("n,T,B,S", [(20, 9, 3, 2), (17, 1, 2, 1), (33, 6, 2, 0), (10, 25, 4, 1), (3, 2, 1, 1)])
def test_forward_matches_oracle(n, T, B, S):
    rng = np.random.default_rng(n * 100 + T)
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    eps = rng.standard_normal((B, T, S, n))
    w_samples, (wE_init, wE_pair, wE_node), w_lognorm = O.natural_lds_inference_general((init, pair), node, eps)
    o, _, _ = run(init, pair, node, eps if S else None, S)
    cpu = lambda x: x.detach().numpy()
    assert rel_err(cpu(o["lognorm"]), w_lognorm) < 1e-11
    assert rel_err(cpu(o["E_node_xx"]), wE_node[0]) < 1e-10 and rel_err(cpu(o["E_node_x"]), wE_node[1]) < 1e-10
    assert rel_err(cpu(o["E_init_xx"]), wE_init[0]) < 1e-10 and rel_err(cpu(o["E_init_x"]), wE_init[1]) < 1e-10
    if T > 1:
        for k, name in enumerate(("E_pair_xx", "E_pair_xxn", "E_pair_xnxn")):
            assert rel_err(cpu(o[name]), wE_pair[k]) < 1e-10, name
    else:
        assert all(float(o[k].abs().max()) == 0.0 for k in ("E_pair_xx", "E_pair_xxn", "E_pair_xnxn"))
    if S:
        assert rel_err(cpu(o["samples"]), w_samples) < 1e-10
    assert int(o["status"]) == 0


def test_inhomogeneous_and_dense_match_oracle():
    rng = np.random.default_rng(7)
    n, T, B, S = 18, 7, 2, 1
    init, pair0 = O.rand_lds_natparam(n, rng)
    steps = [O.rand_lds_natparam(n, rng)[1] for _ in range(T - 1)]
    pair = tuple(np.stack([p[k] for p in steps]) for k in range(4))
    Jd, h, lz = O.rand_node_potentials(B, T, n, rng)
    Q = rng.standard_normal((B, T, n, n)) * 0.1
    Jdense = -0.5 * (np.einsum("btij,btkj->btik", Q, Q) + np.eye(n) * np.log1p(np.exp(rng.standard_normal((B, T, 1, 1)))))
    eps = rng.standard_normal((B, T, S, n))
    for node in ((Jd, h, lz), (Jdense, h, lz)):
        w_samples, (wE_init, wE_pair, wE_node), w_lognorm = O.natural_lds_inference_general((init, pair), node, eps)
        o, _, _ = run(init, pair, node, eps, S, per_t=True)
        cpu = lambda x: x.detach().numpy()
        assert rel_err(cpu(o["lognorm"]), w_lognorm) < 1e-11
        assert rel_err(cpu(o["E_node_xx"]), wE_node[0]) < 1e-10 and rel_err(cpu(o["samples"]), w_samples) < 1e-10
        assert rel_err(cpu(o["Exx_t"][:, :-1]), wE_pair[0]) < 1e-10 and rel_err(cpu(o["Exxn_t"]), wE_pair[1]) < 1e-10
        assert rel_err(cpu(o["Exx_t"][:, 1:]), wE_pair[2]) < 1e-10


def test_vjp_matches_oracle_autograd():
    rng = np.random.default_rng(11)
    n, T, B, S = 19, 6, 2, 1
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    eps = rng.standard_normal((B, T, S, n))
    r = rng.standard_normal
    g_samples, g_lognorm = r((B, T, S, n)), r(B)
    g_stats = ((r((B, n, n)), r((B, n))), (r((B, n, n)), r((B, n, n)), r((B, n, n))), (r((B, T, n)), r((B, T, n))))
    wJ, wh = OT.vjp_node_params((init, pair), node[:2], eps, g_samples, g_stats, g_lognorm)
    o, nJ, nh = run(init, pair, node, eps, S, grad=True)
    total = (o["samples"] * t64(g_samples)).sum() + (o["lognorm"] * t64(g_lognorm)).sum() \
        + (o["E_init_xx"] * t64(g_stats[0][0])).sum() + (o["E_init_x"] * t64(g_stats[0][1])).sum() \
        + (o["E_pair_xx"] * t64(g_stats[1][0])).sum() + (o["E_pair_xxn"] * t64(g_stats[1][1])).sum() \
        + (o["E_pair_xnxn"] * t64(g_stats[1][2])).sum() \
        + (o["E_node_xx"] * t64(g_stats[2][0])).sum() + (o["E_node_x"] * t64(g_stats[2][1])).sum()
    gJ, gh = torch.autograd.grad(total, (nJ, nh))
    assert rel_err(gJ.numpy(), wJ) < 1e-9 and rel_err(gh.numpy(), wh) < 1e-9


def test_message_ports_match_oracle():
    """The three primitives one by one (natural_filter_forward_general, natural_smoother_general, natural_sample_backward) through
    the message modes of the same recursion."""
    rng = np.random.default_rng(21)
    n, T, B, S = 18, 8, 2, 2
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    eps = rng.standard_normal((B, T, S, n))
    (w_pred, w_filt), w_lognorm = O.natural_filter_forward_general(init, pair, node)
    iJ, ih, ilz = (t64(x) for x in init)
    J11, J12, J22, plz = (t64(x) for x in pair)
    o = LL._run(iJ, ih, ilz.reshape(-1), J11, J12, J22, plz, t64(node[0]), t64(node[1]), t64(node[2]), None, 0, False,
                want_stats=False, msgs_out=True)
    assert rel_err(o["J_pred"].numpy(), w_pred[0]) < 1e-11 and rel_err(o["h_pred"].numpy(), w_pred[1]) < 1e-11
    assert rel_err(o["lognorm"].numpy(), w_lognorm) < 1e-11
    J_filt = o["J_pred"] + torch.diag_embed(t64(node[0]))
    h_filt = o["h_pred"] + t64(node[1])
    assert rel_err(J_filt.numpy(), w_filt[0]) < 1e-11 and rel_err(h_filt.numpy(), w_filt[1]) < 1e-11
    wE_init, wE_pair, wE_node = O.natural_smoother_general((w_pred, w_filt), pair)
    w_samples = O.natural_sample_backward_general((w_pred, w_filt), pair, eps)
    z = torch.zeros(B, T, n, dtype=torch.float64)
    o2 = LL._run(-0.5 * torch.eye(n, dtype=torch.float64), z[0, 0], None, J11, J12, J22, None, z, z, None, t64(eps), S, False,
                 from_msgs=(t64(w_filt[0]), t64(w_filt[1])))
    assert rel_err(o2["samples"].numpy(), w_samples) < 1e-10
    assert rel_err(o2["E_node_x"].numpy(), wE_node[1]) < 1e-10 and rel_err(o2["E_node_xx"].numpy(), wE_node[0]) < 1e-10
    assert rel_err(o2["E_init_xx"].numpy(), wE_init[0]) < 1e-10
    for k, name in enumerate(("E_pair_xx", "E_pair_xxn", "E_pair_xnxn")):
        assert rel_err(o2[name].numpy(), wE_pair[k]) < 1e-10, name
