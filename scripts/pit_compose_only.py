"""One launch of svae_lds_compose_chunks (n=10, T=20000, B=1, chunk 128) for an ncu source-level capture."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from oracle import lds_numpy as O
from svae_b200.lds import lds_pit
g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")
rng = np.random.default_rng(0)
n, T = 10, 20000
init, pair = O.rand_lds_natparam(n, rng)
node = O.rand_node_potentials(1, T, n, rng)
lds_pit.compose_chunks(tuple(g(x) for x in pair), tuple(g(x) for x in node), 128)
torch.cuda.synchronize()
