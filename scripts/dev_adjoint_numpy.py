"""Dev tool (not product, not oracle): NumPy prototype of the hand-derived adjoint that
svae_b200/csrc/lds_estep_bwd.cuh implements, checked against oracle/lds_torch.py autograd."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import lds_numpy as O, lds_torch as OT


def forward_save(init, pair, node, eps):
    J0, h0 = -2 * init[0], init[1]
    J11, J12, J22 = -2 * pair[0], -pair[1], -2 * pair[2]
    Jn, hn = node[0], node[1]
    T, n = hn.shape
    S = eps.shape[1]
    P, hp = J0.copy(), h0.copy()
    G, W, m, K, b = [], [], [], [], []
    for t in range(T):
        M = P + np.diag(-2 * Jn[t]) + (J11 if t < T - 1 else 0)
        hf = hp + hn[t]
        L = np.linalg.cholesky(M)
        Kt = np.linalg.inv(L)
        Wt = Kt.T @ Kt
        mt = Wt @ hf
        Gt = -Wt @ J12
        K.append(Kt); W.append(Wt); m.append(mt); G.append(Gt)
        b.append(mt[None] + (Kt.T @ eps[t].T).T)
        P = J22 + J12.T @ Gt
        hp = -J12.T @ mt
    Sig = [None] * T; mu = [None] * T; x = [None] * T
    Sig[T - 1], mu[T - 1], x[T - 1] = W[T - 1], m[T - 1], b[T - 1]
    for t in range(T - 2, -1, -1):
        C = G[t] @ Sig[t + 1]
        Sig[t] = W[t] + C @ G[t].T
        mu[t] = m[t] + G[t] @ mu[t + 1]
        x[t] = b[t] + x[t + 1] @ G[t].T
    return dict(G=G, W=W, m=m, K=K, Sig=Sig, mu=mu, x=x)


def adjoint(sv, eps, g_x, gEinit_xx, gEinit_x, gPxx, gPxxn, gPxnxn, gExx, gEx, g_ln):
    G, W, m, K, Sig, mu, x = (sv[k] for k in ("G", "W", "m", "K", "Sig", "mu", "x"))
    T, n = len(G), G[0].shape[0]
    S = eps.shape[1]
    Sb = np.zeros((n, n)); mub = np.zeros(n); xb = np.zeros((S, n))
    Q = [None] * T; hf1 = [None] * T
    for t in range(T):
        A = np.diag(gExx[t]) + (gEinit_xx if t == 0 else 0) + (gPxx if t < T - 1 else 0) + (gPxnxn if t >= 1 else 0)
        Sb = Sb + A
        mub = mub + gEx[t] + (gEinit_x if t == 0 else 0) + (A + A.T) @ mu[t] \
            + (gPxxn @ mu[t + 1] if t < T - 1 else 0) + (gPxxn.T @ mu[t - 1] if t >= 1 else 0)
        xb = xb + g_x[t]
        Wb, mb = Sb, mub + xb.sum(0)
        if t < T - 1:
            C = G[t] @ Sig[t + 1]
            Cb = gPxxn + Sb @ G[t]
            Gb = Sb.T @ C + Cb @ Sig[t + 1].T + np.outer(mub, mu[t + 1]) + xb.T @ x[t + 1]
            Sb_n, mub_n, xb_n = G[t].T @ Cb, G[t].T @ mub, xb @ G[t]
            inner = Wb @ W[t] + Gb @ G[t].T
        else:
            inner = Wb @ W[t]
        Qt = -W[t] @ inner - np.outer(W[t] @ mb, m[t]) - 0.5 * g_ln * (W[t] + np.outer(m[t], m[t]))
        for s in range(S):
            u = K[t] @ xb[s]
            Xb = -np.tril(np.outer(eps[t, s], u))
            Xb[np.diag_indices(n)] *= 0.5
            Qt = Qt + K[t].T @ Xb @ K[t]
        Q[t] = Qt
        hf1[t] = W[t] @ mb + g_ln * m[t]
        if t < T - 1:
            Sb, mub, xb = Sb_n, mub_n, xb_n
    gJ = np.zeros((T, n)); gh = np.zeros((T, n))
    Pb = hb = None
    for t in range(T - 1, -1, -1):
        if t == T - 1:
            Pb, hb = Q[t], hf1[t]
        else:
            Ghb = G[t] @ hb
            Pb = Q[t] + G[t] @ Pb @ G[t].T - np.outer(Ghb, m[t])
            hb = hf1[t] + Ghb
        gJ[t] = -2 * np.diag(Pb)
        gh[t] = hb
    return gJ, gh


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, T, S in [(3, 5, 2), (2, 1, 1), (4, 2, 1), (6, 9, 3)]:
        init, pair = O.rand_lds_natparam(n, rng)
        node = O.rand_node_potentials(1, T, n, rng)
        eps = rng.standard_normal((1, T, S, n))
        g = dict(g_x=rng.standard_normal((1, T, S, n)), gi_xx=rng.standard_normal((1, n, n)), gi_x=rng.standard_normal((1, n)),
                 gp_xx=rng.standard_normal((1, n, n)), gp_xxn=rng.standard_normal((1, n, n)), gp_xnxn=rng.standard_normal((1, n, n)),
                 gn_xx=rng.standard_normal((1, T, n)), gn_x=rng.standard_normal((1, T, n)), g_ln=rng.standard_normal(1))
        wJ, wh = OT.vjp_node_params((init, pair), node, eps, g["g_x"], ((g["gi_xx"], g["gi_x"]), (g["gp_xx"], g["gp_xxn"], g["gp_xnxn"]),
                                                                    (g["gn_xx"], g["gn_x"])), g["g_ln"])
        sv = forward_save(init, pair, (node[0][0], node[1][0]), eps[0])
        gJ, gh = adjoint(sv, eps[0], g["g_x"][0], g["gi_xx"][0], g["gi_x"][0], g["gp_xx"][0], g["gp_xxn"][0], g["gp_xnxn"][0],
                         g["gn_xx"][0], g["gn_x"][0], g["g_ln"][0])
        print(n, T, S, "err J %.2e h %.2e" % (np.abs(gJ - wJ[0]).max() / np.abs(wJ).max(), np.abs(gh - wh[0]).max() / np.abs(wh).max()))
