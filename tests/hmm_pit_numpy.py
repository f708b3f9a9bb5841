"""NumPy statement of the parallel-in-time HMM E-step pipeline (TEST INFRASTRUCTURE): chunk matrices, boundary
messages, then the sequential oracle (oracle/hmm_numpy.py) on every chunk with the boundary messages folded into its
first and last node potentials.  `stage12` is pluggable: dense NumPy or the host build of the CUDA phase code."""
import ctypes, os, subprocess
import numpy as np
from scipy.special import logsumexp
from oracle import hmm_numpy as HO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stage12_numpy(log_pi, log_A, node, Lc):
    """log chunk matrices by log-space products; log alpha / log beta at the boundaries; total logZ."""
    B, T, K = node.shape
    C = (T - 1 + Lc - 1) // Lc
    la, lb = np.zeros((B, C + 1, K)), np.zeros((B, C + 1, K))
    logPsi = np.zeros((B, C, K, K))
    for b in range(B):
        for c in range(C):
            s = c * Lc
            L = min(Lc, T - 1 - s)
            M = np.where(np.eye(K) > 0, 0.0, -np.inf)
            for t in range(s, s + L):
                step = node[b, t][:, None] + log_A                          # D_t A
                M = logsumexp(M[:, :, None] + step[None], axis=1)
            logPsi[b, c] = M
    la[:, 0] = log_pi - logsumexp(log_pi)
    logZ = np.full(B, logsumexp(log_pi))
    for c in range(C):
        nxt = logsumexp(la[:, c][:, :, None] + logPsi[:, c], axis=1)
        z = logsumexp(nxt, axis=1)
        logZ += z
        la[:, c + 1] = nxt - z[:, None]
    lb[:, C] = node[:, -1]
    logZ += logsumexp(la[:, C] + lb[:, C], axis=1)
    for c in range(C - 1, -1, -1):
        lb[:, c] = logsumexp(logPsi[:, c] + lb[:, c + 1][:, None, :], axis=2)
    return la, lb, logZ


_host = None


def stage12_host(log_pi, log_A, node, Lc):
    global _host
    if _host is None:
        out = os.path.join(ROOT, "tests", "emul", "_build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libhmm_compose_host.so")
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "svae_b200", "csrc"),
                               os.path.join(ROOT, "tests", "emul", "hmm_compose_host.cpp"), "-o", so])
        _host = ctypes.CDLL(so)
    node = np.ascontiguousarray(node, dtype=np.float64)
    B, T, K = node.shape
    C = (T - 1 + Lc - 1) // Lc
    P_ = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lp, lA = np.ascontiguousarray(log_pi, dtype=np.float64), np.ascontiguousarray(log_A, dtype=np.float64)
    Pm, ls = np.zeros((B, C, K, K)), np.zeros((B, C))
    _host.hmm_compose_chunks_host(B, T, K, Lc, P_(lA), P_(node), P_(Pm), P_(ls))
    la, lb, logZ = np.zeros((B, C + 1, K)), np.zeros((B, C + 1, K)), np.zeros(B)
    last = np.ascontiguousarray(node[:, -1])
    _host.hmm_chain_messages_host(B, C, K, P_(lp), P_(Pm), P_(ls), P_(last), P_(la), P_(lb), P_(logZ))
    return la, lb, logZ


def pit_estep(log_pi, log_A, node, Lc, stage12=stage12_numpy):
    """(logZ, (Ez0, Etrans, Ez)) by chunks of Lc steps."""
    B, T, K = node.shape
    C = (T - 1 + Lc - 1) // Lc
    la, lb, logZ = stage12(log_pi, log_A, node, Lc)
    Ez = np.zeros((B, T, K)); Etrans = np.zeros((B, K, K))
    for c in range(C):
        s = c * Lc
        e = min(s + Lc, T - 1)
        sub = node[:, s:e + 1].copy()
        sub[:, 0] += la[:, c]
        sub[:, -1] = lb[:, c + 1]
        _, (_, Et, Ezc) = HO.hmm_estep((np.zeros(K), log_A, sub))
        Etrans += Et
        Ez[:, s:e + 1] = Ezc
    return logZ, (Ez[:, 0], Etrans, Ez)
