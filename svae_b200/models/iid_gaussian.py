"""Plain-VAE special case (no PGM) -- torch mirror of svae/models/iid_gausisan.py:15-46 (diagonal Gaussians;
elementwise, so ordinary torch ops -- SURVEY.md §2 #13)."""
import torch


def unpack_params(natparams):
    neghalfJs, hs = natparams[0], natparams[1]
    Js = -2 * neghalfJs
    return hs / Js, 1. / Js


def sample(natparams, num_samples=None, eps=None, generator=None):
    """iid_gausisan.py:21-26: (T, S, n)."""
    mus, sigmasqs = unpack_params(natparams)
    shape = mus.shape[:-1] + ((num_samples,) if num_samples else ()) + mus.shape[-1:]
    if eps is None:
        eps = torch.randn(shape, dtype=mus.dtype, device=mus.device, generator=generator)
    if num_samples:
        return mus.unsqueeze(-2) + torch.sqrt(sigmasqs).unsqueeze(-2) * eps
    return mus + torch.sqrt(sigmasqs) * eps


def gaussian_vlb(variational_natparams):
    """iid_gausisan.py:29-33: -KL(q || N(0,I))."""
    mus, sigmasqs = unpack_params(variational_natparams)
    return 0.5 * torch.sum(1 + torch.log(sigmasqs) - mus ** 2 - sigmasqs)


def run_inference(prior_natparam, global_natparam, nn_potential, num_samples, eps=None):
    """iid_gausisan.py:42-46 (returns the reference's (samples, (), global_vlb, local_vlb))."""
    return sample(nn_potential, num_samples, eps), (), 0., gaussian_vlb(nn_potential)
