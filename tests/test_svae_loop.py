"""The caller of the path: make_gradfun (svae/svae.py:10-39) around run_inference.

CPU: GMM-SVAE step on the pinwheel recipe (experiments/gmm_svae_synth.py:25-61) and the world_size-2 gloo
data-parallel path (one all-reduce of [statistics | gradients | objective] per step).
GPU: LDS-SVAE steps on synthetic sequences with the fused CUDA E-step (+ VJP) inside the autograd graph."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from svae_b200 import nnet
from svae_b200.svae import make_gradfun
from svae_b200.util import add, scale, flat
from test_gmm import make_pinwheel_data


def _gmm_setup(seed=0, T=40):
    from svae_b200.models import gmm
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    K, N, P = 6, 2, 2
    data = torch.as_tensor(make_pinwheel_data(0.3, 0.05, 5, 40, 0.25, rng))
    pgm_prior = gmm.init_pgm_param(K, N, alpha=0.05 / K, niw_conc=0.5, generator=g)
    pgm = gmm.init_pgm_param(K, N, alpha=1., niw_conc=1., random_scale=3., generator=g)
    recognize, recogn_params = nnet.init_gresnet(P, [(10, torch.tanh), (2 * N, nnet.gaussian_info)], generator=g)
    decode, loglike_params = nnet.init_gresnet(N, [(10, torch.tanh), (2 * P, nnet.gaussian_mean)], generator=g)
    loglike = nnet.make_loglike(decode)
    lab = torch.rand(T, K, dtype=torch.float64, generator=g); lab = lab / lab.sum(-1, keepdim=True)
    run = lambda prior, glob, pots, S: gmm.run_inference(prior, glob, pots, S, label_init=lab[:pots[1].shape[0]],
                                                        generator=torch.Generator().manual_seed(1))
    return run, recognize, loglike, pgm_prior, data, (pgm, loglike_params, recogn_params), T


def test_gmm_svae_step_cpu():
    run, recognize, loglike, pgm_prior, data, params, T = _gmm_setup()
    vals = []
    gradfun = make_gradfun(run, recognize, loglike, pgm_prior, data, T, 4, natgrad_scale=1e2,
                           callback=lambda i, v, p, g: vals.append(v), generator=torch.Generator().manual_seed(0))
    natgrad, ll_grad, rc_grad = gradfun(params, 0)
    assert len(vals) == 1 and np.isfinite(vals[0])
    assert flat(natgrad).shape == flat(params[0]).shape and torch.isfinite(flat(natgrad)).all()
    assert torch.isfinite(flat(ll_grad)).all() and torch.isfinite(flat(rc_grad)).all()
    assert float(flat(rc_grad).abs().max()) > 0


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    run, recognize, loglike, pgm_prior, data, params, T = _gmm_setup()
    half = T // world
    # each rank takes its shard of every minibatch; the reduction must reproduce the single-process step
    shard = lambda pots: tuple(p[rank * half:(rank + 1) * half] for p in pots)
    rec_shard = lambda rp, batch: recognize(rp, batch[rank * half:(rank + 1) * half])
    ll_shard = lambda lp, samples, batch: loglike(lp, samples, batch[rank * half:(rank + 1) * half])
    vals = []
    gradfun = make_gradfun(run, rec_shard, ll_shard, pgm_prior, data, T, 4, natgrad_scale=1e2,
                           callback=lambda i, v, p, g: vals.append(v), generator=torch.Generator().manual_seed(0))
    natgrad, ll_grad, rc_grad = gradfun(params, 0)
    if rank == 0:
        torch.save((flat(natgrad), flat(ll_grad), flat(rc_grad), vals[0]), out)
    dist.destroy_process_group()


def test_data_parallel_allreduce_gloo_world2(tmp_path):
    """Two ranks, each with half of the minibatch: the all-reduced statistics must equal the sum of the
    per-shard statistics (the GMM label factor is per-datapoint, so shards are independent)."""
    out = str(tmp_path / "r0.pt")
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    nat2, ll2, rc2, val2 = torch.load(out)
    # single process: run the two shards by hand and sum
    run, recognize, loglike, pgm_prior, data, params, T = _gmm_setup()
    from svae_b200.util import unbox
    from svae_b200.util import split_into_batches, get_num_datapoints
    batches, nb = split_into_batches(data, T, generator=torch.Generator().manual_seed(0))
    batch = batches[0]
    half = T // 2
    stats_sum = 0
    # the minibatch objective: loglike and local_kl are sums over the datapoints of both shards, global_kl counts once
    from svae_b200.svae import _leaves, _rebuild
    leaves = [l.detach().requires_grad_(True) for l in _leaves((params[1], params[2]))]
    ll_p, rc_p = _rebuild((params[1], params[2]), iter(leaves))
    N = get_num_datapoints(data)
    ll_sum = kl_sum = 0
    for r in range(2):
        sl = slice(r * half, (r + 1) * half)
        pots = recognize(rc_p, batch[sl])
        samples, stats, global_kl, local_kl = run(pgm_prior, unbox(params[0]), pots, 4)
        stats_sum = stats_sum + flat(unbox(stats))
        ll_sum = ll_sum + loglike(ll_p, samples, batch[sl])
        kl_sum = kl_sum + local_kl
    want_val = -(nb * ll_sum - global_kl - nb * kl_sum) / N
    want_grads = torch.autograd.grad(want_val, leaves)
    want_ll, want_rc = _rebuild((params[1], params[2]), iter(want_grads))
    want_nat = -1e2 / N * (flat(pgm_prior) + nb * stats_sum - flat(params[0]))
    assert torch.allclose(nat2, want_nat, rtol=1e-10, atol=1e-12)
    # NN gradients and the objective are SUMS over the shards (not means): same step whatever the world size
    assert torch.allclose(ll2, flat(want_ll), rtol=1e-9, atol=1e-12)
    assert torch.allclose(rc2, flat(want_rc), rtol=1e-9, atol=1e-12)
    want_val = float(want_val.detach())
    assert abs(val2 - want_val) <= 1e-9 * abs(want_val)


@pytest.mark.gpu
def test_lds_svae_training_steps_gpu():
    from svae_b200.models import lds
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    n, p, T, Nseq, Bmb = 4, 6, 30, 64, 16
    # synthetic sequences (svae/lds/synthetic_data.py recipe, vectorised)
    A = torch.randn(n, n, dtype=torch.float64, generator=g); A = A / (torch.linalg.eigvals(A).abs().max() + 0.1)
    C = torch.randn(p, n, dtype=torch.float64, generator=g)
    x = torch.randn(Nseq, n, dtype=torch.float64, generator=g)
    ys = []
    for t in range(T):
        ys.append(x @ C.T + 0.1 * torch.randn(Nseq, p, dtype=torch.float64, generator=g))
        x = x @ A.T + 0.3 * torch.randn(Nseq, n, dtype=torch.float64, generator=g)
    data = torch.stack(ys, 1).to(dev)                                       # (Nseq, T, p)
    prior = lds.make_prior_natparam(n, device=dev)
    pgm = lds.make_prior_natparam(n, scaling=2., device=dev)
    recognize, recogn_params = nnet.init_gresnet(p, [(16, torch.tanh), (2 * n, nnet.gaussian_info)], generator=g, device=dev)
    decode, loglike_params = nnet.init_gresnet(n, [(16, torch.tanh), (2 * p, nnet.gaussian_mean)], generator=g, device=dev)
    loglike = nnet.make_loglike(decode)
    vals = []
    gradfun = make_gradfun(lds.run_inference, recognize, loglike, prior, data, Bmb, 1, natgrad_scale=1.,
                           callback=lambda i, v, p_, g_: vals.append(v), generator=g)
    params = (pgm, loglike_params, recogn_params)
    lr = 1e-3
    for i in range(6):
        natgrad, ll_grad, rc_grad = gradfun(params, i)
        assert torch.isfinite(flat(natgrad)).all() and torch.isfinite(flat(rc_grad)).all()
        params = (add(params[0], scale(-0.05, natgrad)), add(params[1], scale(-lr, ll_grad)), add(params[2], scale(-lr, rc_grad)))
    assert len(vals) == 6 and all(np.isfinite(v) for v in vals)
    assert float(flat(rc_grad).abs().max()) > 0


@pytest.mark.gpu
def test_lds_run_inference_matches_oracle_composition():
    """models/lds.py:45-52 glue: samples, summed global statistics, local KL against the NumPy LDS oracle fed with
    the same local natural parameters."""
    from svae_b200.models import lds
    from svae_b200.distributions import gaussian
    from oracle import lds_numpy as O
    from util import rel_err
    dev = "cuda"
    n, T, B, S = 5, 12, 7, 2
    rng = np.random.default_rng(2)
    prior = lds.make_prior_natparam(n, device=dev)
    glob = lds.make_prior_natparam(n, scaling=1.7, device=dev)
    node = O.rand_node_potentials(B, T, n, rng)[:2]
    eps = rng.standard_normal((B, T, S, n))
    t = lambda x: torch.as_tensor(x, dtype=torch.float64, device=dev)
    samples, stats, global_kl, local_kl = lds.run_inference(prior, glob, (t(node[0]), t(node[1])), S, eps=t(eps))
    eg = lds.lds_prior_expectedstats(glob)
    init = tuple(x.cpu().numpy() for x in gaussian.unpack_dense(eg[0]))
    pair = tuple(x.cpu().numpy() for x in eg[1])
    w_s, (wi, wp, wn), w_ln = O.natural_lds_inference_general(((init[0], init[1], init[2] + init[3]), pair), node + (np.zeros((B, T)),), eps)
    assert rel_err(samples.cpu().numpy(), w_s) < 1e-9
    niw_stats, mniw_stats = stats
    ExxT, Ex, c1, c2 = gaussian.unpack_dense(niw_stats)
    assert rel_err(ExxT.cpu().numpy(), wi[0].sum(0)) < 1e-9 and rel_err(Ex.cpu().numpy(), wi[1].sum(0)) < 1e-9
    assert float(c1) == B and float(c2) == B
    for k in range(3):
        assert rel_err(mniw_stats[k].cpu().numpy(), wp[k].sum(0)) < 1e-9
    assert float(mniw_stats[3]) == B * (T - 1)
    w_kl = (node[0] * wn[0]).sum() + (node[1] * wn[1]).sum() - w_ln.sum()
    assert abs(float(local_kl) - w_kl) < 1e-9 * abs(w_kl)
    assert torch.isfinite(global_kl)


def test_optimizers_on_nested_tuples():
    from svae_b200.optimizers import sgd, adam
    target = (torch.tensor([1., -2.], dtype=torch.float64), (torch.tensor(3., dtype=torch.float64),))
    p0 = (torch.zeros(2, dtype=torch.float64), (torch.zeros((), dtype=torch.float64),))
    gradfun = lambda p, i: (2 * (p[0] - target[0]), (2 * (p[1][0] - target[1][0]),))
    ps = sgd(gradfun, p0, num_iters=400, step_size=0.1)
    pa = adam(gradfun, p0, num_iters=2000, step_size=0.01)
    for q in (ps, pa):
        assert torch.allclose(q[0], target[0], atol=1e-3) and torch.allclose(q[1][0], target[1][0], atol=1e-3)


@pytest.mark.gpu
def test_lds_run_inference_host_equals_run_inference():
    """models/lds.py run_inference for a host-resident minibatch (pinned buffers through LdsHostPipeline): the summed global
    statistics and local_kl equal the device-resident call's; the samples are valid draws (same statistics, own noise)."""
    from svae_b200.models import lds
    n, T, B = 5, 12, 300
    g = torch.Generator().manual_seed(3)
    prior = lds.make_prior_natparam(n, device="cuda")
    glob = lds.make_prior_natparam(n, scaling=1.7, device="cuda")
    J = (-0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64, generator=g))).pin_memory()
    h = torch.randn(B, T, n, dtype=torch.float64, generator=g).pin_memory()
    for S in (0, 1):
        smp_host = torch.zeros((B, T, max(S, 1), n), dtype=torch.float64).pin_memory()
        smp, stats, gkl, lkl = lds.run_inference_host(prior, glob, (J, h), S, smp_host, chunks=[64, 100, 136])
        torch.cuda.synchronize()
        _, wstats, wgkl, wlkl = lds.run_inference(prior, glob, (J.cuda(), h.cuda()), S)
        assert torch.allclose(flat(stats), flat(wstats), rtol=1e-10, atol=1e-10)
        assert abs(float(lkl) - float(wlkl)) < 1e-9 * abs(float(wlkl)) and abs(float(gkl) - float(wgkl)) < 1e-12 * max(1., abs(float(wgkl)))
        if S:
            assert torch.isfinite(smp).all() and float(smp.abs().max()) > 0
