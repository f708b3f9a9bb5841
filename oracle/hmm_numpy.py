"""fp64 NumPy restatement of the reference's HMM local E-step (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/svae/hmm/hmm_inference.py:44-51 (hmm_logZ_python: log-space forward recursion),
/root/reference/svae/hmm/cython_hmm_inference.pyx:93-166 (hmm_logZ and its reverse sweep hmm_logZ_grad,
whose output with g = 1 IS the E-step: hmm_inference.py:65 `hmm_estep_slow = vgrad(hmm_logZ)`), and, for
hmm_estep / hmm_viterbi (hmm_inference.py:21-41,54-63), the third-party routines they delegate to:

    pyhsmm.internals.hmm_messages_interface.{messages_forwards_log, messages_backwards_log,
                                             expected_statistics_log, viterbi}
    -- NOT under /root/reference, no version pinned anywhere in the reference (no requirements file,
       setup.py:5-8 lists no dependencies).  Their published algorithm is the textbook log-space
       alpha/beta recursion, E[z_t] = softmax(alpha_t + beta_t), E[#i->j] = sum_t exp(alpha_t[i] + A[i,j]
       + node[t+1,j] + beta_{t+1}[j] - logZ), and max-product Viterbi with argmax back-pointers.
       Tie-breaking of argmax is not defined in-tree; this restatement takes the FIRST maximal index
       (numpy.argmax) and the tests use tie-free random potentials.

Conventions: natparam = (init_params (K,) = log pi, pair_params (K,K) = log A with row = from-state,
node_params (T,K) log-potentials); all unnormalised log-values are allowed.  A leading batch axis B on the
node parameters is the only addition.  Parity status: pinned against oracle/_ref (compiled
cython_hmm_inference: hmm_logZ, hmm_logZ_grad) in tests/test_oracle_vs_reference.py and tests/golden/hmm_*.npz.
"""
from __future__ import annotations
import numpy as np
from scipy.special import logsumexp


def hmm_logZ(natparam):
    """hmm_inference.py:44-51, batched over node_params (B,T,K).  Returns logZ (B,) and log_alpha (B,T,K)."""
    init_params, pair_params, node_params = natparam
    node_params = np.asarray(node_params, dtype=np.float64)
    B, T, K = node_params.shape
    log_alpha = np.empty((B, T, K))
    log_alpha[:, 0] = init_params + node_params[:, 0]
    for t in range(1, T):
        log_alpha[:, t] = logsumexp(log_alpha[:, t - 1][:, :, None] + pair_params[None], axis=1) + node_params[:, t]
    return logsumexp(log_alpha[:, -1], axis=-1), log_alpha


def hmm_estep(natparam):
    """hmm_inference.py:21-41 (pyhsmm log messages) == vgrad(hmm_logZ) (cython_hmm_inference.pyx:126-166).

    Returns logZ (B,) and (E[z_0] (B,K), E[transition counts] (B,K,K), E[z_t] (B,T,K))."""
    init_params, pair_params, node_params = natparam
    node_params = np.asarray(node_params, dtype=np.float64)
    B, T, K = node_params.shape
    logZ, log_alpha = hmm_logZ(natparam)
    log_beta = np.zeros((B, T, K))
    for t in range(T - 2, -1, -1):
        log_beta[:, t] = logsumexp(pair_params[None] + (node_params[:, t + 1] + log_beta[:, t + 1])[:, None, :], axis=2)
    Ez = np.exp(log_alpha + log_beta - logZ[:, None, None])
    if T > 1:
        lxi = log_alpha[:, :-1, :, None] + pair_params[None, None] + (node_params[:, 1:] + log_beta[:, 1:])[:, :, None, :]
        Etrans = np.exp(lxi - logZ[:, None, None, None]).sum(1)
    else:
        Etrans = np.zeros((B, K, K))
    return logZ, (Ez[:, 0], Etrans, Ez)


def hmm_viterbi(natparam):
    """hmm_inference.py:54-63 (pyhsmm viterbi): most likely state path, int32 (B,T); first-max tie-break."""
    init_params, pair_params, node_params = natparam
    node_params = np.asarray(node_params, dtype=np.float64)
    B, T, K = node_params.shape
    delta = init_params + node_params[:, 0]
    back = np.zeros((B, T, K), dtype=np.int32)
    for t in range(1, T):
        scores = delta[:, :, None] + pair_params[None]            # (B, from, to)
        back[:, t] = np.argmax(scores, axis=1)
        delta = np.max(scores, axis=1) + node_params[:, t]
    path = np.zeros((B, T), dtype=np.int32)
    path[:, -1] = np.argmax(delta, axis=1)
    for t in range(T - 2, -1, -1):
        path[:, t] = back[np.arange(B), t + 1, path[:, t + 1]]
    return path


def rand_hmm_natparam(K, rng):
    """log pi, log A = log Dirichlet(1) rows."""
    pi = rng.dirichlet(np.ones(K))
    A = rng.dirichlet(np.ones(K), size=K)
    return np.log(pi), np.log(A)
