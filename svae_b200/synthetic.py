"""Synthetic E-step problems (host side, NumPy): the benchmark / example workload of SURVEY.md §8(d).

LDS dynamics follow the reference's generator (svae/lds/synthetic_data.py:8-28: A = randn / (max|eig| + 0.1), noise
covariance B B' + n I) converted with svae/lds/gaussian.py:110-143; node potentials mimic the recognition network's
output head (svae/nnet.py:43-47: J = -1/2 softplus(.), h = .); HMM parameters are log Dirichlet(1) rows.
Kept apart from oracle/ (test infrastructure) so that nothing on the product or benchmark path imports the checker.
"""
from __future__ import annotations
import numpy as np


def lds_natparam(n, rng):
    """((-1/2 Sigma0^-1, Sigma0^-1 mu0, logZ0), (-1/2 A'S^-1 A, (S^-1 A)', -1/2 S^-1, -1/2 logdet S))."""
    A = rng.standard_normal((n, n))
    A = A / (np.abs(np.linalg.eigvals(A)).max() + 0.1)
    Bm = rng.standard_normal((n, n))
    S = Bm @ Bm.T + n * np.eye(n)
    C = rng.standard_normal((n, n))
    S0 = C @ C.T + n * np.eye(n)
    mu0 = rng.standard_normal(n)
    h0 = np.linalg.solve(S0, mu0)
    init = (-0.5 * np.linalg.inv(S0), h0, -0.5 * np.dot(mu0, h0) - 0.5 * np.linalg.slogdet(S0)[1])
    SiA = np.linalg.solve(S, A)
    pair = (-0.5 * A.T @ SiA, SiA.T, -0.5 * np.linalg.inv(S), -0.5 * np.linalg.slogdet(S)[1])
    return init, pair


def node_potentials(B, T, n, rng):
    """(J (B,T,n) diagonal natural precisions, h (B,T,n), logZ (B,T) zeros)."""
    J = -0.5 * np.log1p(np.exp(rng.standard_normal((B, T, n))))
    h = rng.standard_normal((B, T, n))
    return J, h, np.zeros((B, T))


def hmm_natparam(K, rng):
    """(log pi (K,), log A (K,K) row = from-state)."""
    pi = rng.dirichlet(np.ones(K))
    A = rng.dirichlet(np.ones(K), size=K)
    return np.log(pi), np.log(A)


def _niw_natparam(S, m, kappa, nu):
    """svae/distributions/niw.py:39-42 (packed dense (N+2, N+2): ((S + kappa m m', kappa m, 0), (0, kappa, 0), (0, 0, nu)))."""
    N = m.shape[0]
    out = np.zeros((N + 2, N + 2))
    out[:N, :N] = S + kappa * np.outer(m, m)
    out[:N, N] = kappa * m
    out[N, N], out[N + 1, N + 1] = kappa, nu
    return out


def _mniw_natparam(nu, S, M, K):
    """svae/distributions/mniw.py:25-31: (K^-1, K^-1 M', S + M K^-1 M', nu)."""
    Kinv = np.linalg.inv(K)
    B = Kinv @ M.T
    return (Kinv, B, S + M @ B, nu)


def gmm_global_natparam(K, N, rng, alpha=1.0, niw_conc=1.0, random_scale=3.0):
    """GMM global natural parameters in the structure of svae/models/gmm.py:33-43: (dirichlet (K,), niw (K,N+2,N+2))."""
    dirichlet = alpha * rng.random(K)
    niw = np.stack([_niw_natparam((N + niw_conc) * np.eye(N), random_scale * rng.standard_normal(N), niw_conc, N + niw_conc)
                    for _ in range(K)])
    return dirichlet, niw


def slds_global_natparam(K, n, rng):
    """SLDS global natural parameters in the structure of svae/models/slds_svae.py:26-75:
    ((dir (K,), mdir (K,K)), [(niw packed (n+2,n+2), mniw (A, B, C, d))] * K), K well-conditioned random modes."""
    dir_nat = 1.0 + rng.random(K)
    mdir_nat = 1.0 + rng.random((K, K)) + 2.0 * np.eye(K)
    lds = []
    for _ in range(K):
        niw = _niw_natparam((n + rng.random()) * np.eye(n), rng.standard_normal(n), 1.0 + rng.random(), n + 4. + rng.random())
        A = rng.standard_normal((n, n))
        A = 0.9 * A / np.abs(np.linalg.eigvals(A)).max()
        lds.append((niw, _mniw_natparam(n + 4. + rng.random(), (n + rng.random()) * np.eye(n), A, (n + rng.random()) * np.eye(n))))
    return (dir_nat, mdir_nat), lds
