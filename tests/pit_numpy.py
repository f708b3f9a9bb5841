"""NumPy statement of the parallel-in-time LDS E-step pipeline (TEST INFRASTRUCTURE).

compose (chunks -> pair potentials), chain (E-step over the chunk boundaries), chunks (independent E-steps with the
chain's prediction message as init and smoothed/predicted as terminal node potential), stitching of the sample paths
with probe lanes.  Every E-step here is the sequential oracle (oracle/lds_numpy.py); `compose` is pluggable so the
host build of the CUDA phase code (tests/emul) can be dropped in.
"""
import ctypes, os, subprocess
import numpy as np
from oracle import lds_numpy as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compose_numpy(pair, node, Lc):
    """Chunk elements by dense linear algebra.  Returns natural-form (cJ11,cJ12,cJ22,clogZ), e1, e2: (B,C,...)."""
    nJ, nh, nlogZ = node
    B, T, n = nh.shape
    C = (T - 1 + Lc - 1) // Lc
    out = [np.zeros((B, C, n, n)) for _ in range(3)] + [np.zeros((B, C)), np.zeros((B, C, n)), np.zeros((B, C, n))]

    def pair_at(b, t):          # homogeneous (n,n), per-time (T-1,n,n) or per-sequence (B,T-1,n,n); info form
        pr = tuple(np.asarray(x) for x in pair)
        if pr[0].ndim == 4:
            pr = tuple(x[b] for x in pr)
        a, bb, cc_, lz = O._pair_at(pr, t)
        return -2 * a, -bb, -2 * cc_, float(lz)

    for b in range(B):
        for c in range(C):
            s = c * Lc
            L = min(Lc, T - 1 - s)
            J11, J12, J22, lz = pair_at(b, s)
            L11 = np.diag(-2 * nJ[b, s]) + J11; L12 = J12.copy(); L22 = J22.copy()
            e1 = nh[b, s].copy(); e2 = np.zeros(n); cc = nlogZ[b, s] + lz
            for t in range(s + 1, s + L):
                J11, J12, J22, lz = pair_at(b, t)
                M = L22 + np.diag(-2 * nJ[b, t]) + J11
                v = e2 + nh[b, t]
                Mi = np.linalg.inv(M)
                cc = cc + nlogZ[b, t] + lz + 0.5 * v @ Mi @ v - 0.5 * np.linalg.slogdet(M)[1]
                L11 = L11 - L12 @ Mi @ L12.T
                e1 = e1 - L12 @ Mi @ v
                L12, L22, e2 = -L12 @ Mi @ J12, J22 - J12.T @ Mi @ J12, -J12.T @ Mi @ v
            out[0][b, c], out[1][b, c], out[2][b, c] = -0.5 * L11, -L12, -0.5 * L22
            out[3][b, c], out[4][b, c], out[5][b, c] = cc, e1, e2
    return tuple(out[:4]), out[4], out[5]


_host = None


def compose_host(pair, node, Lc):
    """The CUDA kernel's phase code compiled for the host (tests/emul/compose_host.cpp)."""
    global _host
    if _host is None:
        out = os.path.join(ROOT, "tests", "emul", "_build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libcompose_host.so")
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "svae_b200", "csrc"),
                               os.path.join(ROOT, "tests", "emul", "compose_host.cpp"), "-o", so])
        _host = ctypes.CDLL(so)
    nJ, nh, nlogZ = (np.ascontiguousarray(x, dtype=np.float64) for x in node)
    B, T, n = nh.shape
    C = (T - 1 + Lc - 1) // Lc
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    pr = [np.ascontiguousarray(x, dtype=np.float64) for x in pair[:3]]
    nd = pr[0].ndim
    st, sb = (0, 0) if nd == 2 else ((n * n, 0) if nd == 3 else (n * n, (T - 1) * n * n))
    lz = np.ascontiguousarray(np.broadcast_to(np.asarray(pair[3], dtype=np.float64), pr[0].shape[:-2] or (1,)))
    pr.append(lz)
    outs = [np.zeros((B, C, n, n)) for _ in range(3)] + [np.zeros((B, C)), np.zeros((B, C, n)), np.zeros((B, C, n))]
    _host.compose_chunks_host.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 4 + [ctypes.c_longlong] * 2 + [ctypes.c_void_p] * 9
    rc = _host.compose_chunks_host(B, T, n, Lc, *[P(x) for x in pr], st, sb, P(nJ), P(nh), P(nlogZ), *[P(x) for x in outs])
    assert rc == 0
    return tuple(outs[:4]), outs[4], outs[5]


def pit_estep(init, pair, node, eps, Lc, compose=compose_numpy):
    """(samples, (E_init, E_pair, E_node), lognorm) by chunks of Lc steps; equals the sequential oracle pointwise."""
    nJ, nh, nlogZ = node
    B, T, n = nh.shape
    S = eps.shape[2]
    C = (T - 1 + Lc - 1) // Lc
    cpair, e1, e2 = compose(pair, node, Lc)
    # boundary chain: states x_{s_0}, ..., x_{s_C = T-1}
    ch = np.zeros((B, C + 1, n)); ch[:, :-1] += e1; ch[:, 1:] += e2; ch[:, -1] += nh[:, -1]
    cJ = np.zeros((B, C + 1, n)); cJ[:, -1] = nJ[:, -1]
    clogZ = np.zeros((B, C + 1)); clogZ[:, -1] = nlogZ[:, -1]
    msgs, lognorm = O.natural_filter_forward_general(init, cpair, (cJ, ch, clogZ))
    _, _, E_node = O.natural_smoother_general(msgs, cpair, diagonal_node_stats=False)
    (Jp, hp), _ = msgs
    hp = hp.copy(); hp[:, 1:] += e2          # eta2 of chunk c-1 rides as node potential of boundary c in the chain
    mu = E_node[1]; Sig = E_node[0] - mu[..., :, None] * mu[..., None, :]
    Js_info = np.linalg.inv(Sig); hs = np.einsum('...ij,...j->...i', Js_info, mu)
    SP = S + 1 + n
    bounds = [(c * Lc, min((c + 1) * Lc, T - 1)) for c in range(C)]
    samples = np.zeros((B, T, S, n)); Exx = np.zeros((B, T, n, n)); Ex = np.zeros((B, T, n))
    pairs = [np.zeros((B, n, n)) for _ in range(3)]
    Exxn = np.zeros((B, T - 1, n, n))
    inhomog = np.asarray(pair[0]).ndim > 2
    xk, Rt = [None] * C, [None] * C
    E_init = None
    for c, (s, e) in enumerate(bounds):
        L = e - s
        sJ = np.zeros((B, L + 1, n, n)); sh = np.zeros((B, L + 1, n)); seps = np.zeros((B, L + 1, SP, n))
        sJ[:, :L] = nJ[:, s:e, :, None] * np.eye(n); sh[:, :L] = nh[:, s:e]
        seps[:, :, :S] = eps[:, s:e + 1]
        sJ[:, L] = -0.5 * Js_info[:, c + 1] - Jp[:, c + 1]
        sh[:, L] = hs[:, c + 1] - hp[:, c + 1]
        seps[:, L, S + 1:] = np.eye(n)
        for b in range(B):
            pr = tuple(np.asarray(x) for x in pair)
            if pr[0].ndim == 4:
                pr = tuple(x[b] for x in pr)
            if pr[0].ndim == 3:
                pr = tuple(x[s:e] for x in pr[:3]) + (np.broadcast_to(pr[3], (T - 1,))[s:e],)
            x, (Ei, Ep, En), _ = O.natural_lds_inference_general(
                ((Jp[b, c], hp[b, c], 0.), pr), (sJ[b:b + 1], sh[b:b + 1], np.zeros((1, L + 1))), seps[b:b + 1])
            if c == 0:
                E_init = E_init or ([], [])
                E_init[0].append(Ei[0][0]); E_init[1].append(Ei[1][0])
            if pr[0].ndim == 3:          # per-time cross moments (the xx blocks are read off Exx below)
                Exxn[b, s:e] = Ep[1][0]
            else:
                for k in range(3):
                    pairs[k][b] += Ep[k][0]
            Exx[b, s:e + 1], Ex[b, s:e + 1] = En[0][0], En[1][0]
            if xk[c] is None:
                xk[c] = np.zeros((B, L + 1, S, n)); Rt[c] = np.zeros((B, L + 1, n, n))
            xk[c][b] = x[0][:, :S]
            D = np.swapaxes(x[0][:, S + 1:] - x[0][:, S:S + 1], -1, -2)
            Rt[c][b] = D @ np.linalg.inv(D[L])
    delta = np.zeros((B, S, n))
    for c in range(C - 1, -1, -1):
        s, e = bounds[c]
        if c < C - 1:
            delta = samples[:, e] - xk[c][:, -1]          # true start of the next chunk minus this chunk's terminal draw
        samples[:, s:e + 1] = xk[c] + np.einsum('btij,bsj->btsi', Rt[c], delta)
    E_init = (np.stack(E_init[0]), np.stack(E_init[1]), np.ones(B), np.ones(B))
    E_pair = (Exx[:, :-1], Exxn, Exx[:, 1:], np.ones((B, T - 1))) if inhomog else (pairs[0], pairs[1], pairs[2], np.full(B, T - 1.0))
    E_nd = (np.diagonal(Exx, axis1=-1, axis2=-2), Ex, np.ones((B, T)))
    return samples, (E_init, E_pair, E_nd), lognorm
