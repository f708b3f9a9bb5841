"""Generates tests/golden/hmm_*.npz from the reference's own compiled cython_hmm_inference (oracle/_ref):
hmm_logZ (pyx:93-121) and hmm_logZ_grad with g = 1 (pyx:126-166) = hmm_estep_slow (hmm_inference.py:65).
Viterbi paths come from the NumPy restatement (pyhsmm is not available; see oracle/hmm_numpy.py header)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref, hmm_numpy as H  # noqa: E402
OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [("hmm_K4_T9", 4, 9, 3, 31), ("hmm_K32_T100", 32, 100, 2, 32), ("hmm_K8_T40", 8, 40, 3, 33), ("hmm_K3_T1", 3, 1, 2, 34)]
for name, K, T, B, seed in CASES:
    rng = np.random.default_rng(seed)
    lp, lA = H.rand_hmm_natparam(K, rng)
    lp, lA = lp + 0.3, lA - 0.2 * rng.random((K, 1))      # unnormalised on purpose
    node = 2.0 * rng.standard_normal((B, T, K))
    outs = [ref.hmm_estep_slow((lp, lA, node[b])) for b in range(B)]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), log_pi=lp, log_A=lA, node=node,
                        logZ=np.array([o[0] for o in outs]), Ez0=np.stack([o[1][0] for o in outs]),
                        Etrans=np.stack([o[1][1] for o in outs]), Ez=np.stack([o[1][2] for o in outs]),
                        viterbi=H.hmm_viterbi((lp, lA, node)))
    print(name, "ok")
