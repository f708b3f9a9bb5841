"""GPU: SLDS composite E-step (HMM forward-backward x per-mode Kalman, svae/models/slds_svae.py:159-175,289-310)
against the NumPy restatement on the BASELINE shape class (small K, d, T so the oracle finishes in seconds)."""
import numpy as np
import pytest
import torch
from oracle import slds_numpy as SO, lds_numpy as O
from util import rel_err

pytestmark = pytest.mark.gpu


def _setup(K=3, n=3, T=12, B=4, seed=0):
    from svae_b200.models import slds_svae as M
    g = torch.Generator().manual_seed(seed)
    glob = M.make_slds_global_natparam(K, n, alpha=2., sticky_bias=1., random=True, generator=g, device="cuda")
    prior = M.make_slds_global_natparam(K, n, device="cuda")
    rng = np.random.default_rng(seed)
    node = O.rand_node_potentials(B, T, n, rng)[:2]
    init_eps = rng.standard_normal((B, T, 1, n))
    return M, prior, glob, node, init_eps


def test_optimize_local_meanfield_matches_oracle():
    M, prior, glob, node, init_eps = _setup()
    t = lambda x: torch.as_tensor(x, dtype=torch.float64, device="cuda")
    (hmm_stats, lds_stats), (hmm_nat, lds_nat), (hmm_vlb, lds_vlb) = M.optimize_local_meanfield(
        glob, (t(node[0]), t(node[1])), init_eps=t(init_eps))
    npg = ((glob[0][0].cpu().numpy(), glob[0][1].cpu().numpy()),
           [(a.cpu().numpy(), tuple(x.cpu().numpy() for x in b)) for a, b in glob[1]])
    w_hmm, (w_init, w_pair, w_node), _, (w_hvlb, w_lvlb), iters = SO.optimize_local_meanfield(npg, node, init_eps)
    assert (1 < np.asarray(iters)).all() and (np.asarray(iters) < 100).all()
    c = lambda x: x.cpu().numpy()
    assert rel_err(c(hmm_stats[2]), w_hmm[2]) < 1e-7 and rel_err(c(hmm_stats[1]), w_hmm[1]) < 1e-7
    assert rel_err(c(lds_stats[0][0]), w_init[0]) < 1e-7
    assert rel_err(c(lds_stats[1][1]), w_pair[1]) < 1e-7
    assert rel_err(c(lds_stats[2][1]), w_node[1]) < 1e-7
    assert rel_err(c(hmm_vlb), w_hvlb) < 1e-7 and rel_err(c(lds_vlb), w_lvlb) < 1e-7


def test_run_inference_differentiable_and_structured():
    M, prior, glob, node, init_eps = _setup(K=4, n=4, T=20, B=6, seed=3)
    from svae_b200.util import flat
    t = lambda x, g=False: torch.as_tensor(x, dtype=torch.float64, device="cuda").requires_grad_(g)
    nJ, nh = t(node[0], True), t(node[1], True)
    samples, stats, global_vlb, local_vlb = M.run_inference(prior, glob, (nJ, nh), 2, init_eps=t(init_eps),
                                                            generator=torch.Generator(device="cuda").manual_seed(0))
    assert samples.shape == (6, 20, 2, 4)
    assert flat(M._tuplify(stats)).shape == flat(M._tuplify(glob)).shape       # svae.py:33-34 lines up
    hmm_g, lds_g = stats
    assert abs(float(hmm_g[0].sum()) - 6) < 1e-9 and abs(float(hmm_g[1].sum()) - 6 * 19) < 1e-7
    (samples.sum() + local_vlb).backward()
    assert torch.isfinite(nJ.grad).all() and torch.isfinite(nh.grad).all() and float(nh.grad.abs().max()) > 0
    assert torch.isfinite(global_vlb)


def test_parallel_in_time_meanfield_matches_sequential():
    """Few long sequences: the chunk-parallel LDS / HMM drivers inside the mean-field loop (same fixed point)."""
    M, prior, glob, node, init_eps = _setup(K=3, n=3, T=300, B=2, seed=5)
    t = lambda x: torch.as_tensor(x, dtype=torch.float64, device="cuda")
    args = (glob, (t(node[0]), t(node[1])))
    a = M.optimize_local_meanfield(*args, init_eps=t(init_eps), parallel_in_time=False)
    b = M.optimize_local_meanfield(*args, init_eps=t(init_eps), parallel_in_time=True)
    c = lambda x: x.cpu().numpy()
    (ah, al), _, (ahv, alv) = a
    (bh, bl), _, (bhv, blv) = b
    assert rel_err(c(bh[2]), c(ah[2])) < 1e-7 and rel_err(c(bh[1]), c(ah[1])) < 1e-7
    assert rel_err(c(bl[1][1]), c(al[1][1])) < 1e-7 and rel_err(c(bl[2][1]), c(al[2][1])) < 1e-7
    assert rel_err(c(bhv), c(ahv)) < 1e-8 and rel_err(c(blv), c(alv)) < 1e-8
    assert M.use_parallel_in_time(1, 20000) and not M.use_parallel_in_time(1024, 200)
