"""GPU: the hand-written reverse-mode kernel (svae_lds_estep_bwd through torch.autograd) against
(1) the committed VJP outputs of the reference's own Cython *_grad functions (tests/golden), and
(2) torch.autograd on the fp64 CPU oracle (oracle/lds_torch.py) for shapes the fixtures do not hold.
Tolerance 1e-8 relative (fp64 both sides; north_star asks 1e-5)."""
import numpy as np
import pytest
import torch
from oracle import lds_numpy as O, lds_torch as OT
from util import golden_cases, load, rel_err, golden_natparams

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def _gpu(x, grad=False):
    t = torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")
    return t.requires_grad_(True) if grad else t


def _vjp_gpu(init, pair, node, eps, g_samples, g_stats, g_lognorm):
    from svae_b200.lds import lds_inference as K
    S = eps.shape[2]
    nJ, nh = _gpu(node[0], True), _gpu(node[1], True)
    nodep = (nJ, nh) + ((_gpu(node[2], True),) if len(node) > 2 else ())
    natparam = (tuple(_gpu(x) for x in init), tuple(_gpu(x) for x in pair))
    samples, (E_init, E_pair, E_node), lognorm = K.natural_lds_inference_general(natparam, nodep, S, _gpu(eps) if S else None)
    (gi_xx, gi_x), gp, (gn_xx, gn_x) = g_stats
    total = (lognorm * _gpu(g_lognorm)).sum() + (E_init[0] * _gpu(gi_xx)).sum() + (E_init[1] * _gpu(gi_x)).sum() \
        + (E_node[0] * _gpu(gn_xx)).sum() + (E_node[1] * _gpu(gn_x)).sum()
    if S:
        total = total + (samples * _gpu(g_samples)).sum()
    for k in range(3):
        total = total + (E_pair[k] * _gpu(gp[k])).sum()
    grads = torch.autograd.grad(total, nodep)
    torch.cuda.synchronize()
    return [g.cpu().numpy() for g in grads]


@pytest.mark.parametrize("path", golden_cases("lds_"))
def test_vjp_golden_from_reference(path):
    d = load(path)
    init, pair, node = golden_natparams(d)
    got = _vjp_gpu(init, pair, node, d["eps"], d["g_samples"],
                   ((d["g_E_init_xx"], d["g_E_init_x"]), (d["g_E_pair_xx"], d["g_E_pair_xxn"], d["g_E_pair_xnxn"]),
                    (d["g_E_node_xx"], d["g_E_node_x"])), d["g_lognorm"])
    assert rel_err(got[0], d["vjp_node_J"]) < RTOL
    assert rel_err(got[1], d["vjp_node_h"]) < RTOL
    assert rel_err(got[2], d["vjp_node_logZ"]) < RTOL


@pytest.mark.parametrize("n,T,B,S,inhomog", [
    (10, 30, 5, 1, False), (16, 12, 3, 2, False), (4, 1, 3, 1, False), (3, 2, 2, 1, False),
    (7, 9, 4, 0, False), (1, 6, 3, 1, False), (5, 8, 3, 2, True), (14, 5, 2, 1, False), (15, 5, 2, 1, False),
])
def test_vjp_vs_torch_oracle(n, T, B, S, inhomog):
    rng = np.random.default_rng(77 * n + T)
    init, pair = O.rand_lds_natparam(n, rng)
    if inhomog:
        pairs = [O.rand_lds_natparam(n, rng)[1] for _ in range(T - 1)]
        pair = tuple(np.stack([p[k] for p in pairs]) for k in range(4))
    node = O.rand_node_potentials(B, T, n, rng)[:2]
    eps = rng.standard_normal((B, T, S, n))
    g_samples = rng.standard_normal((B, T, S, n))
    pshape = (B, T - 1, n, n) if inhomog else (B, n, n)
    g_stats = ((rng.standard_normal((B, n, n)), rng.standard_normal((B, n))),
               tuple(rng.standard_normal(pshape) for _ in range(3)),
               (rng.standard_normal((B, T, n)), rng.standard_normal((B, T, n))))
    g_lognorm = rng.standard_normal(B)
    wJ, wh = OT.vjp_node_params((init, pair), node, eps, g_samples, g_stats, g_lognorm)
    gJ, gh = _vjp_gpu(init, pair, node, eps, g_samples, g_stats, g_lognorm)
    assert rel_err(gJ, wJ) < RTOL and rel_err(gh, wh) < RTOL


def test_local_kl_gradient_like_run_inference():
    """The SVAE use: d/d(node) [ <node, E_node> - lognorm ] + reparameterised samples (models/lds.py:45-52)."""
    from svae_b200.lds import lds_inference as K
    n, T, B, S = 8, 50, 16, 1
    rng = np.random.default_rng(5)
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)[:2]
    eps = rng.standard_normal((B, T, S, n))
    w = rng.standard_normal((B, T, S, n))

    def loss_cpu(Jn, hn):
        t64 = lambda x: torch.as_tensor(x, dtype=torch.float64)
        nat = ((t64(init[0]), t64(init[1]), t64(init[2])), tuple(t64(x) for x in pair))
        samples, (_, _, E_node), lognorm = OT.natural_lds_inference_general(nat, (Jn, hn), t64(eps))
        return ((Jn * E_node[0]).sum() + (hn * E_node[1]).sum() - lognorm.sum()) + (samples * t64(w)).sum()

    Jc = torch.tensor(node[0], requires_grad=True); hc = torch.tensor(node[1], requires_grad=True)
    wJ, wh = torch.autograd.grad(loss_cpu(Jc, hc), (Jc, hc))
    Jg, hg = _gpu(node[0], True), _gpu(node[1], True)
    samples, (_, _, E_node), lognorm = K.natural_lds_inference_general(
        (tuple(_gpu(x) for x in init), tuple(_gpu(x) for x in pair)), (Jg, hg), S, _gpu(eps))
    loss = ((Jg * E_node[0]).sum() + (hg * E_node[1]).sum() - lognorm.sum()) + (samples * _gpu(w)).sum()
    gJ, gh = torch.autograd.grad(loss, (Jg, hg))
    assert rel_err(gJ.cpu().numpy(), wJ.numpy()) < RTOL and rel_err(gh.cpu().numpy(), wh.numpy()) < RTOL
