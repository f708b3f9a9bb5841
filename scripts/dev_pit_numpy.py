"""Dev prototype (NumPy, oracle functions only): parallel-in-time LDS E-step by chunk composition.

 1. compose each chunk's (node, pair) potentials into one pair potential on (x_start, x_end);
 2. run the ordinary E-step on the chain of chunk boundaries (inhomogeneous pair params);
 3. run the ordinary E-step on every chunk independently: init = the chain's prediction message at the chunk
    start, terminal node potential = (smoothed / predicted) at the chunk end, terminal noise chosen so that the
    chunk's last sample IS the chain's sample.
Checks against the sequential oracle.  Not part of the product or the test-suite.
"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle import lds_numpy as O


def compose_chunk(pair_info, nodeJ, nodeh, nodelogZ, pairlogZ):
    """pair_info = (J11,J12,J22) info form, homogeneous; node*(L,n) diag info precision / h / logZ.
    Returns info-form element (L11,L12,L22,e1,e2,c) over steps 0..L (x_0, x_L)."""
    J11, J12, J22 = pair_info
    L, n = nodeh.shape
    L11 = np.diag(nodeJ[0]) + J11; L12 = J12.copy(); L22 = J22.copy()
    e1 = nodeh[0].copy(); e2 = np.zeros(n); c = nodelogZ[0] + pairlogZ
    for t in range(1, L):
        M = L22 + np.diag(nodeJ[t]) + J11
        v = e2 + nodeh[t]
        Mi = np.linalg.inv(M)
        c = c + nodelogZ[t] + pairlogZ + 0.5 * v @ Mi @ v - 0.5 * np.linalg.slogdet(M)[1]
        L11 = L11 - L12 @ Mi @ L12.T
        e1 = e1 - L12 @ Mi @ v
        L12, L22, e2 = -L12 @ Mi @ J12, J22 - J12.T @ Mi @ J12, -J12.T @ Mi @ v
    return L11, L12, L22, e1, e2, c


def pit(init, pair, node, eps, C):
    nJ, nh, nlogZ = node
    B, T, n = nh.shape
    S = eps.shape[2]
    assert (T - 1) % C == 0
    Lc = (T - 1) // C
    J11, J12, J22 = -2 * pair[0], -pair[1], -2 * pair[2]
    # 1. chunk elements
    el = [[compose_chunk((J11, J12, J22), -2 * nJ[b, c * Lc:(c + 1) * Lc], nh[b, c * Lc:(c + 1) * Lc],
                         nlogZ[b, c * Lc:(c + 1) * Lc], pair[3]) for c in range(C)] for b in range(B)]
    st = lambda k: np.array([[el[b][c][k] for c in range(C)] for b in range(B)])
    cpair = (-0.5 * st(0), -st(1), -0.5 * st(2), st(5))
    e1, e2 = st(3), st(4)
    # 2. boundary chain: C+1 states
    ch = np.zeros((B, C + 1, n)); ch[:, :-1] += e1; ch[:, 1:] += e2; ch[:, -1] += nh[:, -1]
    cJ = np.zeros((B, C + 1, n)); cJ[:, -1] = nJ[:, -1]
    clogZ = np.zeros((B, C + 1)); clogZ[:, -1] = nlogZ[:, -1]
    msgs, lognorm = O.natural_filter_forward_general(init, cpair, (cJ, ch, clogZ))
    E_init, E_pair, E_node = O.natural_smoother_general(msgs, cpair, diagonal_node_stats=False)
    (Jp, hp), _ = msgs
    hp = hp.copy(); hp[:, 1:] += e2          # the chain carries eta2 of chunk c-1 as node potential of boundary c
    mu = E_node[1]; Sig = E_node[0] - mu[..., :, None] * mu[..., None, :]
    Js_info = np.linalg.inv(Sig); hs = np.einsum('...ij,...j->...i', Js_info, mu)
    # 3. chunks as independent sequences of length Lc+1; sample lanes: S real + 1 zero-noise probe + n unit probes
    sub_init = (Jp[:, :-1].reshape(B * C, n, n), hp[:, :-1].reshape(B * C, n), 0.)
    SP = S + 1 + n
    sJ = np.zeros((B, C, Lc + 1, n, n)); sh = np.zeros((B, C, Lc + 1, n)); seps = np.zeros((B, C, Lc + 1, SP, n))
    for c in range(C):
        sl = slice(c * Lc, (c + 1) * Lc + 1)
        sJ[:, c, :Lc] = nJ[:, sl, :, None][:, :Lc] * np.eye(n); sh[:, c, :Lc] = nh[:, sl][:, :Lc]
        seps[:, c, :, :S] = eps[:, sl]
    sJ[:, :, Lc] = -0.5 * Js_info[:, 1:] - Jp[:, 1:]
    sh[:, :, Lc] = hs[:, 1:] - hp[:, 1:]
    seps[:, :, Lc, S + 1:] = np.eye(n)
    r = lambda x: x.reshape((B * C,) + x.shape[2:])
    outs = []
    for q in range(B * C):
        init_q = (sub_init[0][q], sub_init[1][q], 0.)
        outs.append(O.natural_lds_inference_general((init_q, pair), (r(sJ)[q:q + 1], r(sh)[q:q + 1], np.zeros((1, Lc + 1))), r(seps)[q:q + 1]))
    xs = np.stack([o[0][0] for o in outs]).reshape(B, C, Lc + 1, SP, n)
    xk = xs[..., :S, :]                                                     # kernel paths
    D = np.swapaxes(xs[..., S + 1:, :] - xs[..., S:S + 1, :], -1, -2)       # (B,C,Lc+1,n,n): column j = probe j
    Rt = D @ np.linalg.inv(D[:, :, Lc])[:, :, None]                         # G_{t->e} for every t
    delta = np.zeros((B, C, S, n))
    for c in range(C - 2, -1, -1):                                          # affine suffix recursion over chunks
        x_true_start = xk[:, c + 1, 0] + np.einsum('bij,bsj->bsi', Rt[:, c + 1, 0], delta[:, c + 1])
        delta[:, c] = x_true_start - xk[:, c, Lc]
    samples = xk + np.einsum('bctij,bcsj->bctsi', Rt, delta)
    Exx = np.stack([o[1][2][0][0] for o in outs]).reshape(B, C, Lc + 1, n, n)
    Ex = np.stack([o[1][2][1][0] for o in outs]).reshape(B, C, Lc + 1, n)
    pairs = [np.stack([o[1][1][k][0] for o in outs]).reshape(B, C, n, n).sum(1) for k in range(3)]
    full = lambda a: np.concatenate([a[:, :, :Lc].reshape((B, C * Lc) + a.shape[3:]), a[:, -1, Lc:]], 1)
    E_init_o = (np.stack([outs[b * C][1][0][0][0] for b in range(B)]), np.stack([outs[b * C][1][0][1][0] for b in range(B)]))
    return full(samples), (E_init_o, pairs, (np.diagonal(full(Exx), axis1=-1, axis2=-2), full(Ex))), lognorm


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    n, T, B, S, C = 4, 25, 2, 2, 4
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    node = (node[0], node[1], 0.3 * rng.standard_normal((B, T)))
    eps = rng.standard_normal((B, T, S, n))
    w_samples, (wE_init, wE_pair, wE_node), w_lognorm = O.natural_lds_inference_general((init, pair), node, eps)
    samples, (E_init, E_pair, E_node), lognorm = pit(init, pair, node, eps, C)
    err = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    print("lognorm", err(lognorm, w_lognorm))
    print("E_node", err(E_node[0], wE_node[0]), err(E_node[1], wE_node[1]))
    print("E_init", err(E_init[0], wE_init[0]), err(E_init[1], wE_init[1]))
    for k in range(3):
        print("E_pair", k, err(E_pair[k], wE_pair[k]))
    print("samples (pointwise, same noise)", err(samples, w_samples))
