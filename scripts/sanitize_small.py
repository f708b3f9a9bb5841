"""Dev helper for compute-sanitizer: one small launch of every kernel (forward, VJP, HMM, Viterbi, reduce)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lds_numpy as O, hmm_numpy as H
from svae_b200.lds import lds_inference as K
from svae_b200.hmm import hmm_inference as HK
rng = np.random.default_rng(0)
g = lambda x, grad=False: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda").requires_grad_(grad)
for n, T, B, S in [(10, 12, 5, 1), (16, 6, 3, 2), (3, 5, 7, 1)]:
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)[:2]
    nJ, nh = g(node[0], True), g(node[1], True)
    samples, (Ei, Ep, En), ln = K.natural_lds_inference_general((tuple(g(x) for x in init), tuple(g(x) for x in pair)), (nJ, nh), S)
    (samples.sum() + ln.sum() + En[0].sum() + Ep[1].sum()).backward()
    o = K.lds_estep_forward(tuple(g(x) for x in init), tuple(g(x) for x in pair), (nJ.detach(), nh.detach()), S)
    K.lds_reduce_global_stats(o)
lp, lA = H.rand_hmm_natparam(7, rng)
nat = (g(lp), g(lA), g(rng.standard_normal((5, 9, 7))))
HK.hmm_estep(nat); HK.hmm_viterbi(nat)
torch.cuda.synchronize()
print("sanitize run ok")
