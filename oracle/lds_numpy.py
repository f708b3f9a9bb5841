"""fp64 NumPy restatement of the reference's LDS local E-step (TEST INFRASTRUCTURE ONLY).

Follows, function by function, the Cholesky forms in /root/reference/svae/lds/gaussian.py:37-121 and
the drivers in /root/reference/svae/lds/lds_inference.py:86-202, with ONE addition: every routine
carries a leading batch axis B of independent sequences (the reference handles one sequence per
call; SURVEY.md "No batch-of-sequences axis exists").

Conventions are the reference's: natural parameters are (-1/2 J, h[, logZ]); pair parameters are
(J11, J12, J22, logZ) = (-1/2 A'S^-1 A, (S^-1 A)', -1/2 S^-1, -1/2 logdet S)  (gaussian.py:130-143);
node parameters are (J_node, h_node[, logZ_node]) with J_node diagonal (T,n) or dense (T,n,n)
(lds_inference.py:65-80).

Shapes here: init = (J0 (n,n), h0 (n), logZ0) shared by the batch; pair = shared homogeneous
(n,n) x3 + scalar, or inhomogeneous (T-1,n,n)/(T-1,), or per-sequence (B,T-1,n,n)/(B,T-1);
node = (B,T,n) / (B,T,n,n), (B,T,n), (B,T).

Parity status: pinned against oracle/_ref (the reference's Cython compiled here) by
tests/test_oracle_vs_reference.py and against tests/golden/*.npz generated from it.
"""
from __future__ import annotations
import numpy as np


def _T(x):
    return np.swapaxes(x, -1, -2)


def _solve_tri(L, b, trans=False):
    """solve_triangular(L, b[, 'T']) batched; b is (...,n) or (...,n,k)."""
    vec = b.ndim == L.ndim - 1
    bb = b[..., None] if vec else b
    A = _T(L) if trans else L
    out = np.linalg.solve(A, bb)  # L is triangular; exact enough for an oracle, n <= 16
    return out[..., 0] if vec else out


# ---------------------------------------------------------------------------------------------
# per-step Gaussian ops in natural form                                   (svae/lds/gaussian.py)
# ---------------------------------------------------------------------------------------------

def natural_condition_on_general(J, h, Jo, ho, logZo):
    """gaussian.py:46-49.  Jo diagonal (...,n) or dense (...,n,n)."""
    if Jo.ndim == h.ndim:
        Jo = Jo[..., None] * np.eye(Jo.shape[-1])
    return (J + Jo, h + ho), logZo


def natural_predict(J, h, J11, J12, J22, logZ):
    """gaussian.py:52-66 (Cholesky form; what cython_gaussian_grads.pxd:38-82 computes)."""
    J, J11, J12, J22 = -2 * J, -2 * J11, -J12, -2 * J22
    L = np.linalg.cholesky(J + J11)
    v = _solve_tri(L, h)
    lognorm = 0.5 * np.sum(v * v, -1) - np.sum(np.log(np.diagonal(L, axis1=-1, axis2=-2)), -1)
    v2 = _solve_tri(L, v, trans=True)
    h_predict = -np.einsum('...ji,...j->...i', J12, v2)
    temp = _solve_tri(L, np.broadcast_to(J12, L.shape))
    J_predict = J22 - _T(temp) @ temp
    return (-0.5 * J_predict, h_predict), lognorm + logZ


def natural_lognorm(J, h):
    """gaussian.py:80-84."""
    J = -2 * J
    L = np.linalg.cholesky(J)
    v = _solve_tri(L, h)
    return 0.5 * np.sum(v * v, -1) - np.sum(np.log(np.diagonal(L, axis1=-1, axis2=-2)), -1)


def info_to_mean(J, h):
    """gaussian.py:117-121."""
    Sigma = np.linalg.inv(J)
    mu = np.einsum('...ij,...j->...i', Sigma, h)
    return mu, Sigma


def natural_rts_backward_step(next_smooth, next_pred, filtered, pair_param):
    """gaussian.py:87-106."""
    (Jns, hns, mun), (Jnp, hnp), (Jf, hf) = next_smooth, next_pred, filtered
    J11, J12, J22 = pair_param[:3]
    Jns, Jnp, Jf, J11, J12, J22 = -2 * Jns, -2 * Jnp, -2 * Jf, -2 * J11, -J12, -2 * J22

    J11, J12, J22 = Jf + J11, J12, Jns - Jnp + J22
    L = np.linalg.cholesky(J22)
    temp = _solve_tri(L, np.broadcast_to(_T(J12), L.shape))
    Js = J11 - _T(temp) @ temp
    hs = hf - np.einsum('...ji,...j->...i', temp, _solve_tri(L, hns - hnp))

    mu, sigma = info_to_mean(Js, hs)
    # -solve_posdef_from_cholesky(L, J12.T sigma) + outer(mun, mu)
    rhs = _T(np.broadcast_to(J12, L.shape)) @ sigma
    ExnxT = -_solve_tri(L, _solve_tri(L, rhs), trans=True) + mun[..., :, None] * mu[..., None, :]
    ExxT = sigma + mu[..., :, None] * mu[..., None, :]
    return -0.5 * Js, hs, (mu, ExxT, ExnxT)


def natural_condition_on(J, h, y, Jxx, Jxy):
    """gaussian.py:37-43 with y (...,S,n): returns J_cond (...,n,n), h_cond (...,S,n)."""
    return J + Jxx, h[..., None, :] + np.einsum('...ij,...sj->...si', Jxy, y)


def natural_sample(J, h, eps):
    """gaussian.py:69-77 with the noise injected: J (...,n,n), h (...,S,n), eps (...,S,n)."""
    J = -2 * J
    L = np.linalg.cholesky(J)
    noise = _solve_tri(L, _T(eps), trans=True)                       # (...,n,S)
    mean = _solve_tri(L, _solve_tri(L, _T(h)), trans=True)           # (...,n,S)
    return _T(mean + noise)


# ---------------------------------------------------------------------------------------------
# parameter bookkeeping                                          (svae/lds/lds_inference.py:44-80)
# ---------------------------------------------------------------------------------------------

def _canonical_init_params(init_params):
    """lds_inference.py:61-62."""
    return init_params[0], init_params[1], sum(init_params[2:])


def _canonical_node_params(node_params):
    """lds_inference.py:65-80, plus the leading batch axis."""
    node_params = tuple(np.asarray(p, dtype=np.float64) for p in node_params)
    if len(node_params) not in (2, 3):
        raise ValueError
    J, h = node_params[:2]
    if h.ndim != 3 or J.ndim not in (3, 4):
        raise ValueError
    B, T, n = h.shape
    logZ = node_params[2] if len(node_params) == 3 else np.zeros((B, T))
    if J.shape not in ((B, T, n), (B, T, n, n)) or logZ.shape != (B, T):
        raise ValueError
    return J, h, logZ


def _pair_at(pair_params, t):
    """_repeat_param (lds_inference.py:50-59): homogeneous (n,n) or time-indexed (...,T-1,n,n)."""
    J11, J12, J22, logZ = pair_params
    J11, J12, J22 = (np.asarray(x, dtype=np.float64) for x in (J11, J12, J22))
    logZ = np.asarray(logZ, dtype=np.float64)
    if J11.ndim == 2:
        return J11, J12, J22, logZ
    return J11[..., t, :, :], J12[..., t, :, :], J22[..., t, :, :], logZ[..., t]


def pair_is_homogeneous(pair_params):
    return np.asarray(pair_params[0]).ndim == 2


# ---------------------------------------------------------------------------------------------
# filter / smoother / sampler                                   (svae/lds/lds_inference.py:86-178)
# ---------------------------------------------------------------------------------------------

def natural_filter_forward_general(init_params, pair_params, node_params):
    """lds_inference.py:86-106 == cython_lds_inference.pyx:28-90.

    Returns ((J_pred (B,T,n,n), h_pred (B,T,n)), (J_filt, h_filt)) in natural form and lognorm (B,).
    """
    J0, h0, logZ0 = _canonical_init_params(init_params)
    Jn, hn, logZn = _canonical_node_params(node_params)
    B, T, n = hn.shape
    J = np.broadcast_to(np.asarray(J0, dtype=np.float64), (B, n, n))
    h = np.broadcast_to(np.asarray(h0, dtype=np.float64), (B, n))
    lognorm = np.zeros(B) + logZ0
    Jp, hp, Jf, hf = [], [], [], []
    for t in range(T):
        Jp.append(J); hp.append(h)
        (J, h), term = natural_condition_on_general(J, h, Jn[:, t], hn[:, t], logZn[:, t])
        lognorm = lognorm + term
        Jf.append(J); hf.append(h)
        if t < T - 1:
            (J, h), term = natural_predict(J, h, *_pair_at(pair_params, t))
            lognorm = lognorm + term
    lognorm = lognorm + natural_lognorm(J, h)
    st = lambda xs: np.stack(xs, 1)
    return ((st(Jp), st(hp)), (st(Jf), st(hf))), lognorm


def natural_smoother_general(forward_messages, pair_params, diagonal_node_stats=True):
    """lds_inference.py:127-178 == cython_lds_inference.pyx:149-210.

    Returns E_init = (ExxT_0, Ex_0, 1., 1.), E_pair = (sum_t ExxT_t, sum_t E[x_t x_{t+1}'],
    sum_t ExxT_{t+1}, T-1) per sequence when homogeneous, per-t (B,T-1,...) otherwise,
    E_node = (diag ExxT (B,T,n) | ExxT (B,T,n,n), Ex (B,T,n), ones (B,T)).
    """
    (Jp, hp), (Jf, hf) = forward_messages
    B, T, n = hf.shape
    mu, Sigma = info_to_mean(-2 * Jf[:, -1], hf[:, -1])                  # natural_to_mean
    Js, hs = Jf[:, -1], hf[:, -1]
    stats = [(mu, Sigma + mu[:, :, None] * mu[:, None, :], None)]
    for t in range(T - 2, -1, -1):
        Js, hs, (mu_t, ExxT, ExnxT) = natural_rts_backward_step(
            (Js, hs, mu), (Jp[:, t + 1], hp[:, t + 1]), (Jf[:, t], hf[:, t]), _pair_at(pair_params, t))
        mu = mu_t
        stats.insert(0, (mu, ExxT, ExnxT))
    Ex = np.stack([s[0] for s in stats], 1)
    ExxT = np.stack([s[1] for s in stats], 1)
    ones = np.ones(B)
    E_init = (ExxT[:, 0], Ex[:, 0], ones, ones)
    if T > 1:
        ExxnT = np.stack([_T(s[2]) for s in stats[:-1]], 1)              # make_pair_stats: ExnxT.T
        E_pair = (ExxT[:, :-1], ExxnT, ExxT[:, 1:], np.ones((B, T - 1)))
    else:
        z = np.zeros((B, 0, n, n))
        E_pair = (z, z, z, np.ones((B, 0)))
    if pair_is_homogeneous(pair_params):
        E_pair = tuple(x.sum(1) for x in E_pair)
    node_xx = np.diagonal(ExxT, axis1=-1, axis2=-2) if diagonal_node_stats else ExxT
    E_node = (node_xx, Ex, np.ones((B, T)))
    return E_init, E_pair, E_node


def natural_sample_backward_general(forward_messages, pair_params, eps):
    """lds_inference.py:109-124 == cython_lds_inference.pyx:310-355, noise injected.

    eps (B,T,S,n): eps[:, t] is the standard-normal draw consumed at time t.  (The reference draws
    randn(T,S,N) and consumes it in REVERSE time order, pyx:333: its draw k goes to time T-1-k.)
    Returns samples (B,T,S,n).
    """
    _, (Jf, hf) = forward_messages
    B, T, n = hf.shape
    S = eps.shape[2]
    x = natural_sample(Jf[:, -1], np.broadcast_to(hf[:, -1][:, None, :], (B, S, n)), eps[:, -1])
    out = [x]
    for t in range(T - 2, -1, -1):
        J11, J12 = _pair_at(pair_params, t)[:2]
        Jc, hc = natural_condition_on(Jf[:, t], hf[:, t], x, J11, J12)
        x = natural_sample(Jc, hc, eps[:, t])
        out.insert(0, x)
    return np.stack(out, 1)


def natural_lds_inference_general(natparam, node_params, eps):
    """lds_inference.py:185-202: (samples, (E_init, E_pair, E_node), lognorm)."""
    init_params, pair_params = natparam
    msgs, lognorm = natural_filter_forward_general(init_params, pair_params, node_params)
    diag = np.asarray(node_params[0]).ndim == 3
    stats = natural_smoother_general(msgs, pair_params, diagonal_node_stats=diag)
    samples = natural_sample_backward_general(msgs, pair_params, eps)
    return samples, stats, lognorm


def natural_lds_estep_general(natparam, node_params):
    """lds_inference.py:223-237: (lognorm, stats)."""
    init_params, pair_params = natparam
    msgs, lognorm = natural_filter_forward_general(init_params, pair_params, node_params)
    diag = np.asarray(node_params[0]).ndim == 3
    return lognorm, natural_smoother_general(msgs, pair_params, diagonal_node_stats=diag)


# ---------------------------------------------------------------------------------------------
# converting + synthetic problems        (svae/lds/gaussian.py:110-143, svae/lds/synthetic_data.py:8-28)
# ---------------------------------------------------------------------------------------------

def mean_to_natural(mu, sigma):
    """gaussian.py:110-114."""
    neghalfJ = -0.5 * np.linalg.inv(sigma)
    h = np.linalg.solve(sigma, mu)
    logZ = -0.5 * np.dot(mu, h) - 0.5 * np.linalg.slogdet(sigma)[1]
    return neghalfJ, h, logZ


def pair_mean_to_natural(A, sigma):
    """gaussian.py:130-143."""
    temp = np.linalg.solve(sigma, A)
    Jxx = -0.5 * _T(A) @ temp
    Jxy = _T(temp)
    Jyy = -0.5 * np.linalg.inv(sigma)
    logZ = -0.5 * np.linalg.slogdet(sigma)[1]
    return Jxx, Jxy, Jyy, logZ


def rand_lds_natparam(n, rng):
    """Synthetic dynamics per synthetic_data.py:8-28 and SURVEY.md §8d:
    A = randn/(max|eig|+0.1), Sigma = BB'+n I; init = mean_to_natural(randn, rand_psd+n I)."""
    A = rng.standard_normal((n, n))
    A = A / (np.abs(np.linalg.eigvals(A)).max() + 0.1)
    Bm = rng.standard_normal((n, n))
    sigma = Bm @ Bm.T + n * np.eye(n)
    C = rng.standard_normal((n, n))
    sigma0 = C @ C.T + n * np.eye(n)
    mu0 = rng.standard_normal(n)
    return mean_to_natural(mu0, sigma0), pair_mean_to_natural(A, sigma)


def rand_node_potentials(B, T, n, rng):
    """Mimics nnet.gaussian_info (svae/nnet.py:43-47): J = -1/2 softplus(randn), h = randn."""
    J = -0.5 * np.log1p(np.exp(rng.standard_normal((B, T, n))))
    h = rng.standard_normal((B, T, n))
    return J, h, np.zeros((B, T))
