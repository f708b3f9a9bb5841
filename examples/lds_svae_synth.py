#!/usr/bin/env python
"""LDS-SVAE on synthetic sequences, end to end on one B200 -- the experiment the reference's svae/models/lds.py:70-78
helpers were written for, in the shape of experiments/gmm_svae_synth.py:25-61: recognition + decoder gresnets
(PyTorch), natural-gradient SVI on the NIW x MNIW global factor, and the fused CUDA E-step (+ its reverse-mode
kernel) inside the autograd graph.

    python examples/lds_svae_synth.py [--iters 200] [--batch 256]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svae_b200 import nnet
from svae_b200.svae import make_gradfun
from svae_b200.models import lds
from svae_b200.optimizers import adam


def make_data(num_seq, T, n, p, gen, device):
    """svae/lds/synthetic_data.py:8-48 recipe, vectorised over sequences."""
    A = torch.randn(n, n, dtype=torch.float64, generator=gen); A = A / (torch.linalg.eigvals(A).abs().max() + 0.1)
    C = torch.randn(p, n, dtype=torch.float64, generator=gen)
    x = torch.randn(num_seq, n, dtype=torch.float64, generator=gen)
    ys = []
    for _ in range(T):
        ys.append(x @ C.T + 0.1 * torch.randn(num_seq, p, dtype=torch.float64, generator=gen))
        x = x @ A.T + 0.3 * torch.randn(num_seq, n, dtype=torch.float64, generator=gen)
    return torch.stack(ys, 1).to(device)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("-T", type=int, default=50)
    ap.add_argument("-n", type=int, default=8)
    ap.add_argument("-p", type=int, default=12)
    ap.add_argument("--pgm-on-device", dest="pgm_on_host", action="store_false")
    a = ap.parse_args()
    dev = "cuda"
    gen = torch.Generator().manual_seed(0)
    data = make_data(a.seq, a.T, a.n, a.p, gen, dev)
    # the global (PGM) block stays on the host: its NIW/MNIW maps are two n x n matrices, microseconds on a CPU
    # core and launch-latency-bound on a GPU; only the ~4n^2 local natural parameters cross PCIe per step
    pgm_dev = "cpu" if a.pgm_on_host else dev
    prior = lds.make_prior_natparam(a.n, device=pgm_dev)
    pgm = lds.make_prior_natparam(a.n, scaling=2., device=pgm_dev)
    recognize, recogn_params = nnet.init_gresnet(a.p, [(32, torch.tanh), (2 * a.n, nnet.gaussian_info)], generator=gen, device=dev)
    decode, loglike_params = nnet.init_gresnet(a.n, [(32, torch.tanh), (2 * a.p, nnet.gaussian_mean)], generator=gen, device=dev)
    loglike = nnet.make_loglike(decode)
    log = []
    gradfun = make_gradfun(lds.run_inference, recognize, loglike, prior, data, a.batch, 1, natgrad_scale=1e2,
                           callback=lambda i, v, p_, g_: log.append(v), generator=gen)
    t0 = time.perf_counter()
    adam(gradfun, (pgm, loglike_params, recogn_params), num_iters=a.iters, step_size=1e-3)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("LDS-SVAE n=%d p=%d T=%d minibatch=%d: -ELBO/datapoint %.4f -> %.4f in %d iterations, %.1f ms/iteration"
          % (a.n, a.p, a.T, a.batch, log[0], sum(log[-5:]) / 5, a.iters, dt / a.iters * 1e3))


if __name__ == "__main__":
    main()
