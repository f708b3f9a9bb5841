"""One pass of each parallel-in-time driver (for an ncu launch list): LDS n=10 T=100000 B=1 S=1, HMM K=32 T=100000 B=1."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from oracle import lds_numpy as O, hmm_numpy as HO
from svae_b200.lds import lds_pit
from svae_b200.hmm import hmm_pit

g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")
rng = np.random.default_rng(0)
n, T = 10, 100000
init, pair = O.rand_lds_natparam(n, rng)
node = O.rand_node_potentials(1, T, n, rng)
nat, nd = (tuple(g(x) for x in init), tuple(g(x) for x in pair)), tuple(g(x) for x in node)
eps = torch.randn((1, T, 1, n), dtype=torch.float64, device="cuda")
lds_pit.natural_lds_inference_parallel(nat, nd, 1, eps, chunk_len=128)
lp, lA = HO.rand_hmm_natparam(32, rng)
hmm_pit.hmm_estep_parallel((g(lp), g(lA), torch.randn((1, T, 32), dtype=torch.float64, device="cuda")))
torch.cuda.synchronize()
print("done")
