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


def viterbi_stage12_host(log_pi, log_A, node, Lc):
    """Max-plus chunk matrices and optimal boundary states by the host build of the CUDA phase code."""
    stage12_host(log_pi, log_A, node[:, :2] if node.shape[1] > 2 else node, 1)      # make sure the library is built
    node = np.ascontiguousarray(node, dtype=np.float64)
    B, T, K = node.shape
    C = (T - 1 + Lc - 1) // Lc
    P_ = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lp, lA = np.ascontiguousarray(log_pi, dtype=np.float64), np.ascontiguousarray(log_A, dtype=np.float64)
    V = np.zeros((B, C, K, K))
    _host.hmm_compose_chunks_maxplus_host(B, T, K, Lc, P_(lA), P_(node), P_(V))
    delta = np.zeros((B, C + 1, K)); zb = np.zeros((B, C + 1), dtype=np.int32)
    last = np.ascontiguousarray(node[:, -1])
    _host.hmm_chain_viterbi_host(B, C, K, P_(lp), P_(V), P_(last), P_(delta), P_(zb))
    return V, zb


def viterbi_stage12_numpy(log_pi, log_A, node, Lc):
    B, T, K = node.shape
    C = (T - 1 + Lc - 1) // Lc
    V = np.zeros((B, C, K, K))
    for b in range(B):
        for c in range(C):
            s = c * Lc
            L = min(Lc, T - 1 - s)
            M = np.where(np.eye(K) > 0, 0.0, -1e300)
            for t in range(s, s + L):
                M = np.max(M[:, :, None] + (node[b, t][:, None] + log_A)[None], axis=1)
            V[b, c] = M
    delta = np.zeros((B, C + 1, K)); delta[:, 0] = log_pi
    for c in range(C):
        delta[:, c + 1] = np.max(delta[:, c][:, :, None] + V[:, c], axis=1)
    zb = np.zeros((B, C + 1), dtype=np.int32)
    zb[:, C] = np.argmax(delta[:, C] + node[:, -1], axis=1)
    for c in range(C - 1, -1, -1):
        zb[:, c] = np.argmax(delta[:, c] + V[np.arange(B), c, :, zb[:, c + 1]], axis=1)
    return V, zb


def pit_viterbi(log_pi, log_A, node, Lc, stage12=viterbi_stage12_numpy):
    """Most likely path by chunks: boundary states from the max-plus chain, then the sequential oracle on every chunk
    with both end states pinned."""
    B, T, K = node.shape
    C = (T - 1 + Lc - 1) // Lc
    _, zb = stage12(log_pi, log_A, node, Lc)
    path = np.zeros((B, T), dtype=np.int32)
    never = -1e300
    for c in range(C):
        s = c * Lc
        e = min(s + Lc, T - 1)
        sub = node[:, s:e + 1].copy()
        pin0 = np.full((B, K), never); pin0[np.arange(B), zb[:, c]] = 0.0
        pin1 = np.full((B, K), never); pin1[np.arange(B), zb[:, c + 1]] = 0.0
        sub[:, 0] += pin0
        sub[:, -1] = pin1
        path[:, s:e + 1] = HO.hmm_viterbi((np.zeros(K), log_A, sub))
    return path
