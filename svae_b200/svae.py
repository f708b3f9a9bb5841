"""make_gradfun -- torch mirror of the reference's svae/svae.py:10-39 (the SVI step around the E-step).

Same curried signature and the same returned gradient triple
    gradfun(params, i) -> (pgm_natgrad, loglike_grad, recogn_grad),  params = (pgm_params, loglike_params, recogn_params)
so an optimiser loop written against the reference runs unchanged; torch.autograd plays the role of
autograd.value_and_grad over (loglike_params, recogn_params) (svae.py:28-30), the natural gradient of the PGM
parameters is formed from the expected statistics the E-step returns (svae.py:31-34, no backprop into
pgm_params), and with torch.distributed initialised the statistics / gradients of a sharded minibatch are
summed with ONE all-reduce per step (a sum, not a mean: the shards partition the minibatch).
"""
from __future__ import annotations
import torch
from .util import flatten, get_num_datapoints, split_into_batches, unbox

callback = lambda i, val, params, grad: print('{}: {}'.format(i, val))
flat = lambda struct: flatten(struct)[0]


def _leaves(struct):
    if isinstance(struct, (tuple, list)):
        out = []
        for s in struct:
            out += _leaves(s)
        return out
    return [struct]


def _rebuild(struct, it):
    if isinstance(struct, (tuple, list)):
        return tuple(_rebuild(s, it) for s in struct)
    return next(it)


def make_gradfun(run_inference, recognize, loglike, pgm_prior, data, batch_size, num_samples,
                 natgrad_scale=1., callback=callback, generator=None):
    """svae.py:10-39."""
    _, unflat = flatten(pgm_prior)
    num_datapoints = get_num_datapoints(data)
    data_batches, num_batches = split_into_batches(data, batch_size, generator=generator)
    get_batch = lambda i: data_batches[i % num_batches]
    dist = torch.distributed if torch.distributed.is_available() and torch.distributed.is_initialized() else None

    def gradfun(params, i):
        pgm_params, loglike_params, recogn_params = params
        batch = get_batch(i)
        nn_leaves = [l.detach().requires_grad_(True) for l in _leaves((loglike_params, recogn_params))]
        ll_p, rc_p = _rebuild((loglike_params, recogn_params), iter(nn_leaves))
        # mc_elbo (svae.py:19-24)
        nn_potentials = recognize(rc_p, batch)
        samples, stats, global_kl, local_kl = run_inference(pgm_prior, unbox(pgm_params), nn_potentials, num_samples)
        elbo = (num_batches * loglike(ll_p, samples, batch) - global_kl - num_batches * local_kl) / num_datapoints
        val = -elbo
        grads = torch.autograd.grad(val, nn_leaves, allow_unused=True)
        grads = [torch.zeros_like(l) if g is None else g for g, l in zip(grads, nn_leaves)]
        stats_flat = flat(unbox(stats))
        pgm_flat = flat(unbox(pgm_params))
        if dist is not None and dist.get_world_size() > 1:
            # data-parallel minibatch: one all-reduce of [expected statistics | NN gradients | objective]
            buf = torch.cat([stats_flat] + [g.reshape(-1) for g in grads] + [val.detach().reshape(1)])
            dist.all_reduce(buf)
            k = stats_flat.numel()
            stats_flat = buf[:k]
            off = k
            for j, g in enumerate(grads):
                grads[j] = buf[off:off + g.numel()].reshape(g.shape)
                off += g.numel()
            # loglike and local_kl are sums over the datapoints of a shard, so the minibatch gradient / objective is
            # the SUM over ranks; global_kl (no NN dependence) sits in every rank's val and is counted once
            val = buf[-1] - (dist.get_world_size() - 1) * torch.as_tensor(global_kl).detach().to(buf) / num_datapoints
        loglike_grad, recogn_grad = _rebuild((loglike_params, recogn_params), iter(grads))
        # natural gradient of the PGM parameters (svae.py:31-34)
        # (the PGM block may live on the host -- its exp-family maps are K tiny matrices -- while the statistics
        #  come back from the device: one small copy, the same payload as the all-reduce)
        pgm_natgrad = -natgrad_scale / num_datapoints * \
            (flat(pgm_prior).to(pgm_flat.device) + num_batches * stats_flat.to(pgm_flat.device) - pgm_flat)
        grad = unflat(pgm_natgrad), loglike_grad, recogn_grad
        if callback:
            callback(i, float(val.detach()), params, grad)
        return grad

    return gradfun
