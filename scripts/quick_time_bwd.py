import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lds_numpy as O
from svae_b200.lds import lds_inference as K
n, T, B, S = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (10, 200, 4096, 1))]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
rng = np.random.default_rng(0)
init, pair = O.rand_lds_natparam(n, rng)
g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")
gi, gp = tuple(g(x) for x in init), tuple(g(x) for x in pair)
nJ = -0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64, device="cuda"))
nh = torch.randn(B, T, n, dtype=torch.float64, device="cuda")
eps = torch.randn(B, T, S, n, dtype=torch.float64, device="cuda")
w = torch.randn(B, T, S, n, dtype=torch.float64, device="cuda")
gl = -torch.ones(B, dtype=torch.float64, device="cuda")
def fwd(): return K.lds_estep_forward(gi, gp, (nJ, nh), S, eps, save_for_bwd=True)
def bwd(o): K.lds_estep_backward(o, dict(samples=w, E_node_xx=nJ, E_node_x=nh, lognorm=gl))
for _ in range(2): bwd(fwd())
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
for _ in range(iters):
    ev[0].record(); o = fwd(); ev[1].record(); bwd(o); ev[2].record(); torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
print("n=%d T=%d B=%d S=%d: fwd(save) %.3f ms  bwd %.3f ms  -> %.3e seq/s fwd+bwd" % (n, T, B, S, tf / iters, tb / iters, B / ((tf + tb) / iters) * 1e3))
