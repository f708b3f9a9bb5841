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
