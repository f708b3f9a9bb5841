"""fp64 NumPy restatement of the reference's SLDS composite local E-step (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/svae/models/slds_svae.py:80-175,229-243 (lds_meanfield, get_var_lds_local_natparam,
hmm_meanfield, get_arhmm_local_nodeparams, optimize_local_meanfield, get_global_stats), composed from
oracle/lds_numpy.py and oracle/hmm_numpy.py, with /root/reference/svae/distributions/mniw.py:42-55 and
niw.py:15-24 for the per-mode expected parameters.  Batched over B (stop when every sequence converged).
The random initialisation (slds_svae.py:203-226: one sample path of a 0.9-random-walk LDS) takes injected noise.
Parity status: the reference's Python layer cannot run under Python 3 and holds no vectors for this path --
PARITY UNPINNED beyond the pinned LDS / HMM pieces it is composed of.
"""
import numpy as np
from scipy.special import digamma
from . import lds_numpy as L, hmm_numpy as H
from .gmm_numpy import niw_expectedstats, unpack_dense, dirichlet_expectedstats


def mniw_expectedstats(natparam, fudge=1e-8):
    A, B, C, d = natparam
    K = np.linalg.inv(A); K = (K + K.T) / 2
    M = (K @ B).T
    S = C - M @ B
    nu, m = d, M.shape[0]
    sym = lambda X: (X + X.T) / 2
    E_Sigmainv = nu * sym(np.linalg.inv(S)) + fudge * np.eye(S.shape[0])
    E_Sigmainv_A = nu * np.linalg.solve(S, M)
    E_AT_Sigmainv_A = m * K + nu * sym(M.T @ np.linalg.solve(S, M)) + fudge * np.eye(K.shape[0])
    E_logdet = digamma((nu - np.arange(m)) / 2.).sum() + m * np.log(2) - np.linalg.slogdet(S)[1]
    return (-0.5 * E_AT_Sigmainv_A, E_Sigmainv_A.T, -0.5 * E_Sigmainv, 0.5 * E_logdet)


def all_local(lds_global):
    inits = [unpack_dense(niw_expectedstats(a)) for a, _ in lds_global]
    pairs = [mniw_expectedstats(b) for _, b in lds_global]
    return tuple(np.stack(x) for x in zip(*inits)), tuple(np.stack(x) for x in zip(*pairs))


def arhmm_node(lds_global, init_stats, pair_stats):
    (J, h, a, b), (J11, J12, J22, logZ) = all_local(lds_global)
    init_node = np.einsum('bij,kij->bk', init_stats[0], J) + np.einsum('bi,ki->bk', init_stats[1], h) + (a + b)[None]
    rest = np.einsum('btij,kij->btk', pair_stats[0], J11) + np.einsum('btij,kij->btk', pair_stats[1], J12) \
        + np.einsum('btij,kij->btk', pair_stats[2], J22) + logZ[None, None]
    return np.concatenate((init_node[:, None], rest), 1)


def lds_estep_perseq(init_param, pair_params, node):
    """per-sequence init: loop over B (the NumPy LDS oracle shares init across its batch)."""
    B = node[0].shape[0]
    outs = []
    for b in range(B):
        init = (init_param[0][b], init_param[1][b], init_param[2][b] + init_param[3][b])
        pair = tuple(p[b] for p in pair_params)
        ln, stats = L.natural_lds_estep_general((init, pair), tuple(x[b:b + 1] for x in node) + (np.zeros((1, node[1].shape[1])),))
        outs.append((ln[0], stats))
    lognorm = np.array([o[0] for o in outs])
    cat = lambda f: np.concatenate([f(o[1]) for o in outs])
    E_init = (cat(lambda s: s[0][0]), cat(lambda s: s[0][1]))
    E_pair = (cat(lambda s: s[1][0]), cat(lambda s: s[1][1]), cat(lambda s: s[1][2]))
    E_node = (cat(lambda s: s[2][0]), cat(lambda s: s[2][1]))
    return lognorm, E_init, E_pair, E_node


def optimize_local_meanfield(global_natparam, node, init_eps, tol=1e-2, max_iter=100):
    (dir_nat, mdir_nat), lds_global = global_natparam
    B, T, n = node[1].shape
    A = 0.9 * np.eye(n)
    rw = ((-0.5 * np.eye(n), np.zeros(n), 0.), (-0.5 * A.T @ A, A.T, -0.5 * np.eye(n), 0.))
    msgs, _ = L.natural_filter_forward_general(rw[0], rw[1], node + (np.zeros((B, T)),))
    x = L.natural_sample_backward_general(msgs, rw[1], init_eps)[:, :, 0]
    out = lambda u, v: u[..., :, None] * v[..., None, :]
    init_stats = (out(x[:, 0], x[:, 0]), x[:, 0])
    pair_stats = (out(x[:, :-1], x[:, :-1]), out(x[:, :-1], x[:, 1:]), out(x[:, 1:], x[:, 1:]))
    log_pi, log_A = dirichlet_expectedstats(dir_nat), dirichlet_expectedstats(mdir_nat)
    dense_init, dense_pair = all_local(lds_global)
    # every sequence stops on its own bound (frozen afterwards): B sequences == B reference calls (slds_svae.py:159-175)
    active = np.ones(B, dtype=bool)
    sel = lambda new, old: np.where(active.reshape((-1,) + (1,) * (np.ndim(new) - 1)), new, old)
    tsel = lambda new, old: tuple(tsel(a, b) if isinstance(a, tuple) else sel(a, b) for a, b in zip(new, old))
    cur, prev, iters = None, None, np.zeros(B, dtype=int)
    for it in range(1, max_iter + 1):
        hnode = arhmm_node(lds_global, init_stats, pair_stats)
        hmm_vlb, hmm_stats = H.hmm_estep((log_pi, log_A, hnode))
        Ez = hmm_stats[2]
        init_param = tuple(np.tensordot(Ez[:, 0], p, 1) for p in dense_init)
        pair_params = tuple(np.tensordot(Ez[:, 1:], p, 1) for p in dense_pair)
        lds_vlb, init_stats_n, pair_stats_n, E_node = lds_estep_perseq(init_param, pair_params, node)
        new = (tuple(hmm_stats), (init_stats_n, pair_stats_n, E_node), (init_param, pair_params), (hmm_vlb, lds_vlb))
        cur = new if cur is None else tsel(new, cur)
        iters[active] = it
        vlb = cur[3][0] + cur[3][1]
        if prev is not None:
            active = active & ~(np.abs(vlb - prev) < tol)
            if not active.any():
                break
        prev = vlb
        init_stats, pair_stats = cur[1][0], cur[1][1]
    hmm_stats, lds_all, nat, vlbs = cur
    return hmm_stats, lds_all, nat, vlbs, (int(iters[0]) if B == 1 else iters)
