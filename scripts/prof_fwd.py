"""Dev helper for ncu: ONE launch of the fused LDS E-step at a given shape (after one warm-up launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svae_b200 import synthetic
from svae_b200.lds import lds_inference as K

n, T, B, S = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (10, 200, 4096, 1))]
save = len(sys.argv) > 5 and sys.argv[5] == "save"
torch.manual_seed(0)
g = lambda x: torch.as_tensor(x, dtype=torch.float64, device="cuda")
import numpy as np
from svae_b200.synthetic import lds_natparam as rand_lds_natparam
init, pair = rand_lds_natparam(n, np.random.default_rng(0))
gi, gp = tuple(g(x) for x in init), tuple(g(x) for x in pair)
nJ = -0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64, device="cuda"))
nh = torch.randn(B, T, n, dtype=torch.float64, device="cuda")
eps = torch.randn(B, T, S, n, dtype=torch.float64, device="cuda") if S else None
for _ in range(2):
    K.lds_estep_forward(gi, gp, (nJ, nh), S, eps, save_for_bwd=save)
torch.cuda.synchronize()
