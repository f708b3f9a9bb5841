#!/usr/bin/env python
"""GMM-SVAE on the 2-D pinwheel -- the reference's experiments/gmm_svae_synth.py:25-61 on one B200.

Same model and hyper-parameters as the reference script: K = 15 mixture components on N = 2 latent dimensions,
P = 2 observed, 5 x 100 pinwheel points, 40-40 tanh Gaussian resnets for recognition (gaussian_info head) and decoder
(gaussian_mean head), sparsifying Dirichlet / NIW prior, SGD (momentum 0.9, step 1e-3) on the network parameters and
natural gradient (scale 1e4) on the GMM global factor, minibatches of 50 points, one sample per point.

Per step: recognition gresnet (PyTorch GEMMs + the fused output head csrc/nnet_heads.cu) -> svae_gmm_meanfield (the whole
label <-> Gaussian coordinate ascent of svae/models/gmm.py:90-110 in ONE launch) -> differentiable evaluation at the optimum
-> decoder log-likelihood -> torch.autograd for the network gradients, expected statistics for the natural gradient
(svae/svae.py:10-39).  Plotting (make_plotter_2d) is out of scope; the callback logs the objective.

    python examples/gmm_svae_synth.py [--iters 1000] [--device cuda]
"""
import argparse, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svae_b200 import nnet
from svae_b200.svae import make_gradfun
from svae_b200.models.gmm import run_inference, init_pgm_param, make_encoder_decoder
from svae_b200.optimizers import sgd


def make_pinwheel_data(radial_std, tangential_std, num_classes, num_per_class, rate, generator):
    """experiments/gmm_svae_synth.py:11-23."""
    rads = torch.linspace(0, 2 * math.pi, num_classes + 1, dtype=torch.float64)[:-1]
    features = torch.randn(num_classes * num_per_class, 2, dtype=torch.float64, generator=generator) \
        * torch.tensor([radial_std, tangential_std], dtype=torch.float64)
    features[:, 0] += 1.
    labels = torch.arange(num_classes).repeat_interleave(num_per_class)
    angles = rads[labels] + rate * torch.exp(features[:, 0])
    rotations = torch.stack([torch.cos(angles), -torch.sin(angles), torch.sin(angles), torch.cos(angles)]).T.reshape(-1, 2, 2)
    out = 10 * torch.einsum('ti,tij->tj', features, rotations)
    return out[torch.randperm(out.shape[0], generator=generator)]


def main(iters=1000, device="cuda", seed=1, quiet=False):
    gen = torch.Generator().manual_seed(seed)
    num_clusters, samples_per_cluster = 5, 100
    K, N, P = 15, 2, 2
    data = make_pinwheel_data(0.3, 0.05, num_clusters, samples_per_cluster, 0.25, gen).to(device)
    # prior: sparsifying but otherwise generic (gmm_svae_synth.py:39)
    pgm_prior_params = init_pgm_param(K, N, alpha=0.05 / K, niw_conc=0.5, generator=gen, device=device)
    recognize, recogn_params = nnet.init_gresnet(P, [(40, torch.tanh), (40, torch.tanh), (2 * N, nnet.gaussian_info)],
                                                 generator=gen, device=device)
    decode, loglike_params = nnet.init_gresnet(N, [(40, torch.tanh), (40, torch.tanh), (2 * P, nnet.gaussian_mean)],
                                               generator=gen, device=device)
    loglike = nnet.make_loglike(decode)
    pgm_params = init_pgm_param(K, N, alpha=1., niw_conc=1., random_scale=3., generator=gen, device=device)
    params = (pgm_params, loglike_params, recogn_params)
    encode_mean, decode_mean = make_encoder_decoder(recognize, decode)
    log = []

    def callback(i, val, params, grad):
        log.append(val)
        if not quiet and i % 100 == 0:
            print("iteration %4d   -ELBO / datapoint %.4f" % (i, val))
    dev_gen = torch.Generator(device=device).manual_seed(seed) if device != "cpu" else gen
    inference = lambda prior, glob, pots, S: run_inference(prior, glob, pots, S, generator=dev_gen)
    gradfun = make_gradfun(inference, recognize, loglike, pgm_prior_params, data, batch_size=50, num_samples=1,
                           natgrad_scale=1e4, callback=callback, generator=gen)
    t0 = time.perf_counter()
    params = sgd(gradfun, params, num_iters=iters, step_size=1e-3)
    if device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if not quiet:
        z = encode_mean(data, params[0], params[2])
        print("GMM-SVAE K=%d: -ELBO/datapoint %.3f -> %.3f in %d iterations, %.2f ms/iteration; latent means span [%.2f, %.2f]"
              % (K, sum(log[:10]) / 10, sum(log[-10:]) / 10, iters, dt / iters * 1e3, float(z.min()), float(z.max())))
    return log, params


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args()
    main(a.iters, a.device)
