import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from svae_b200.models import slds_svae as M, gmm
from svae_b200.distributions import dirichlet, niw
from oracle import lds_numpy as O
K, n, T, B = 8, 10, 200, 1024
g = torch.Generator().manual_seed(0)
glob = M.make_slds_global_natparam(K, n, alpha=2., sticky_bias=1., random=True, generator=g, device="cuda")
prior = M.make_slds_global_natparam(K, n, device="cuda")
nJ = -0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64, device="cuda"))
nh = torch.randn(B, T, n, dtype=torch.float64, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(1)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        (hs, ls), nat, (hv, lv) = M.optimize_local_meanfield(glob, (nJ, nh), generator=gen)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    samples, stats, gv, lvlb = M.run_inference(prior, glob, (nJ, nh), 1, generator=gen)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("SLDS K=%d d=%d T=%d B=%d: optimize_local_meanfield %.1f ms; full run_inference %.1f ms -> %.0f seq/s" % (K, n, T, B, (t1 - t0) * 1e3, (t2 - t1) * 1e3, B / (t2 - t1)))
# GMM fixed point kernel
Kg, N, Tg, Bg = 10, 2, 50, 256
pgm = gmm.init_pgm_param(Kg, N, alpha=0.05 / Kg, niw_conc=0.5, random_scale=3.0, generator=g)
lg = dirichlet.expectedstats(pgm[0].cuda()); gg = niw.expectedstats(pgm[1].cuda())
gJ = -0.5 * torch.nn.functional.softplus(torch.randn(Bg, Tg, N, dtype=torch.float64, device="cuda")); gh = 2 * torch.randn(Bg, Tg, N, dtype=torch.float64, device="cuda")
init = torch.rand(Bg, Tg, Kg, dtype=torch.float64, device="cuda"); init = init / init.sum(-1, keepdim=True)
for _ in range(3): out = gmm.meanfield_fixed_point_cuda(lg, gg, gJ, gh, init)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out = gmm.meanfield_fixed_point_cuda(lg, gg, gJ, gh, init)
e1.record(); torch.cuda.synchronize()
print("GMM K=%d N=%d T=%d x %d minibatches: fixed point %.3f ms/launch (mean %.1f sweeps)" % (Kg, N, Tg, Bg, e0.elapsed_time(e1) / 10, float(out[1].double().mean())))
