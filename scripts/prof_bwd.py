"""Dev helper for ncu: forward(save) + VJP launches of the LDS E-step at a given shape (training cotangents)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from svae_b200.synthetic import lds_natparam
from svae_b200.lds import lds_inference as K
n, T, B, S = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (10, 200, 4096, 1))]
g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")
init, pair = lds_natparam(n, np.random.default_rng(0))
gi, gp = tuple(g(x) for x in init), tuple(g(x) for x in pair)
nJ = -0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64, device="cuda"))
nh = torch.randn(B, T, n, dtype=torch.float64, device="cuda")
eps = torch.randn(B, T, S, n, dtype=torch.float64, device="cuda")
w = torch.randn(B, T, S, n, dtype=torch.float64, device="cuda")
gl = -torch.ones(B, dtype=torch.float64, device="cuda")
for _ in range(2):
    o = K.lds_estep_forward(gi, gp, (nJ, nh), S, eps, save_for_bwd=True)
    K.lds_estep_backward(o, dict(samples=w, E_node_xx=nJ, E_node_x=nh, lognorm=gl))
torch.cuda.synchronize()
