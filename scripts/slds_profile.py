"""Dev helper: where run_inference of the SLDS spends its time at the BASELINE shape (host-side wall clock, synchronised)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from svae_b200 import synthetic as SY
from svae_b200.models import slds_svae as M
K, n, T, B = 8, 10, 200, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")
(dn, mn), lds_np = SY.slds_global_natparam(K, n, np.random.default_rng(0))
glob = ((g(dn), g(mn)), [(g(a), tuple(g(x) for x in b)) for a, b in lds_np])
nJ = -0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64, device="cuda"))
nh = torch.randn(B, T, n, dtype=torch.float64, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(1)
def timed(f, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
for rep in range(3):
    _, t_all = timed(M.run_inference, glob, glob, (nJ, nh), 1, generator=gen)
    dense, t_dense = timed(M.get_all_lds_local_natparams, glob[1])
    with torch.no_grad():
        r, t_opt = timed(M.optimize_local_meanfield, glob, (nJ, nh), generator=gen, dense=dense)
    _, t_vlb = timed(M.slds_prior_vlb, glob, glob)
    (hs, ls), (hn, ln), _ = r
    _, t_gs = timed(M.get_global_stats, hs, ls[:2])
    print("B=%d run_inference %.1f ms | per-mode maps %.1f | optimize_local_meanfield %.1f | prior vlb %.1f | global stats %.1f"
          % (B, t_all, t_dense, t_opt, t_vlb, t_gs))
import torch.autograd.profiler as P
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    M.run_inference(glob, glob, (nJ, nh), 1, generator=gen); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
