"""Dev helper for ncu: run_inference of the SLDS at the BASELINE shape (K=8, d=10, T=200), twice (second call = steady state)."""
import sys, os
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
for _ in range(2):
    M.run_inference(glob, glob, (nJ, nh), 1, generator=gen)
torch.cuda.synchronize()
