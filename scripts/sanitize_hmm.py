"""Dev helper for compute-sanitizer: small launches of every HMM kernel added in round 2's second session -- the tensor-core
sweeps (four-warp default; SVAE_HMM_TC_WARPS=1 selects the one-warp sweep) with the transition-count kernel, the wide kernels
(K > 32: preparation, E-step, Viterbi), ragged batches, masks, log-normaliser-only launches, T = 1, 2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from svae_b200 import synthetic as SY
from svae_b200.hmm import hmm_inference as HK
rng = np.random.default_rng(0)
g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")
for K, T, B in [(32, 9, 11), (17, 5, 8), (25, 1, 3), (32, 2, 9), (20, 12, 1), (33, 7, 3), (64, 1, 2), (70, 10, 2), (8, 6, 4)]:
    lp, lA = SY.hmm_natparam(K, rng)
    node = g(3.0 * rng.standard_normal((B, T, K)))
    nat = (g(lp), g(lA), node)
    logZ, (E0, Et, Ez) = HK.hmm_estep(nat)
    HK.hmm_estep_forward(nat[0], nat[1], node, logz_only=True)
    active = torch.ones(B, dtype=torch.int32, device="cuda"); active[::2] = 0
    HK.hmm_estep_forward(nat[0], nat[1], node, active=active, out=(logZ.clone(), E0.clone(), Et.clone(), Ez.clone()))
    HK.hmm_viterbi(nat)
    assert bool(torch.isfinite(logZ).all()) and abs(float(Ez.sum()) - B * T) < 1e-6 * B * T
torch.cuda.synchronize()
print("sanitize hmm run ok")
