"""Dev helper for compute-sanitizer: small launches of the parallel-in-time kernels (LDS / HMM compose, HMM chain), the
message ports, dense node potentials (forward + VJP) and the cp.async pair-parameter staging of the INHOMOG path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lds_numpy as O, hmm_numpy as H
from svae_b200.lds import lds_inference as K, lds_pit
from svae_b200.hmm import hmm_pit
rng = np.random.default_rng(0)
g = lambda x, grad=False: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda").requires_grad_(grad)
for n, T, B, S, Lc in [(10, 30, 3, 1, 7), (16, 20, 2, 1, 6), (3, 14, 2, 2, 4)]:
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)[:2]
    nat = (tuple(g(x) for x in init), tuple(g(x) for x in pair))
    lds_pit.natural_lds_inference_parallel(nat, tuple(g(x) for x in node), S, chunk_len=Lc)
    msgs, _ = K.natural_filter_forward_general(nat[0], nat[1], tuple(g(x) for x in node))
    K.natural_smoother_general(msgs, nat[1]); K.natural_sample_backward(msgs, nat[1], S)
    R = rng.standard_normal((B, T, n, n)) / np.sqrt(n)
    Jd = g(-0.5 * (R @ np.swapaxes(R, -1, -2) + 0.1 * np.eye(n)), True)
    smp, (Ei, Ep, En), ln = K.natural_lds_inference_general(nat, (Jd, g(node[1], True)), S)
    (smp.sum() + ln.sum() + En[0].sum()).backward()
    prs = [O.rand_lds_natparam(n, rng)[1] for _ in range(B * (T - 1))]
    ipair = tuple(g(np.stack([p[k] for p in prs]).reshape((B, T - 1) + np.shape(prs[0][k]))) for k in range(4))
    lds_pit.natural_lds_inference_parallel((nat[0], ipair), tuple(g(x) for x in node), S, chunk_len=Lc)
for Kk, T, B, Lc in [(32, 40, 2, 7), (5, 23, 3, 4)]:
    lp, lA = H.rand_hmm_natparam(Kk, rng)
    hmm_pit.hmm_estep_parallel((g(lp), g(lA), g(rng.standard_normal((B, T, Kk)))), chunk_len=Lc)
torch.cuda.synchronize()
print("sanitize pit run ok")
