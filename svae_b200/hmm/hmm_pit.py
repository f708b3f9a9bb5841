"""Parallel-in-time HMM E-step for the small-batch / long-T regime.

Same contract and numbers as hmm_estep (svae/hmm/hmm_inference.py:21-41), from three wide launches:
  1. svae_hmm_compose_chunks -- each chunk's  prod_t diag(exp(node_t)) A  as one scaled K x K matrix (B x C warps);
  2. svae_hmm_chain_messages -- forward / backward messages at the chunk boundaries and the total logZ (B warps,
     one vector-matrix product per chunk and direction);
  3. svae_hmm_estep on all B x C chunks as independent sequences: forward message added to the chunk's first node
     potential, backward message in place of its last one, log pi = 0 -- so every chunk's forward-backward returns
     the exact posterior marginals and expected transition counts of its steps.
hmm_logZ_parallel is differentiable: the VJP of logZ IS the E-step (cython_hmm_inference.pyx:126-166), which the
same three launches return.
"""
from __future__ import annotations
import math
import torch
from .. import _lib
from . import hmm_inference as H
from ..lds.lds_pit import default_chunk_len


def hmm_estep_parallel(natparam, chunk_len=None):
    """(logZ, (E[z_0], E[transition counts], E[z_t]))."""
    lp, lA, node, batched, dev = H._canon(natparam)
    B, T, K = node.shape
    Lc = int(chunk_len) if chunk_len else default_chunk_len(T)
    if T < 4 or Lc >= T - 1:
        return H.hmm_estep(natparam)
    L = _lib.lib()
    C = L.svae_lds_compose_num_chunks(T, Lc)
    new = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)
    P_, ls = new(B, C, K, K), new(B, C)
    ptr, st = _lib.ptr, _lib.stream_ptr(dev)
    _lib.check(L.svae_hmm_compose_chunks(B, T, K, Lc, ptr(lA), ptr(node), ptr(P_), ptr(ls), st), "svae_hmm_compose_chunks")
    la, lb, logZ = new(B, C + 1, K), new(B, C + 1, K), new(B)
    last = node[:, -1].contiguous()
    _lib.check(L.svae_hmm_chain_messages(B, C, K, ptr(lp), ptr(P_), ptr(ls), ptr(last), ptr(la), ptr(lb), ptr(logZ), st),
               "svae_hmm_chain_messages")
    nfull = (T - 1) // Lc
    tail = (T - 1) - nfull * Lc
    Ez = new(B, T, K)
    Etrans = torch.zeros((B, K, K), dtype=torch.float64, device=dev)
    zero_pi = torch.zeros(K, dtype=torch.float64, device=dev)
    Ez0 = None
    for c0, cnt, Ls in [(0, nfull, Lc)] + ([(nfull, 1, tail)] if tail else []):
        t0, t1 = c0 * Lc, c0 * Lc + cnt * Ls
        sub = new(B, cnt, Ls + 1, K)
        sub[:, :, :Ls] = node[:, t0:t1].reshape(B, cnt, Ls, K)
        sub[:, :, 0] += la[:, c0:c0 + cnt]
        sub[:, :, Ls] = lb[:, c0 + 1:c0 + cnt + 1]
        _, e0, et, ez = H.hmm_estep_forward(zero_pi, lA, sub.reshape(B * cnt, Ls + 1, K))
        Etrans += et.reshape(B, cnt, K, K).sum(1)
        ez = ez.reshape(B, cnt, Ls + 1, K)
        Ez[:, t0:t1] = ez[:, :, :Ls].reshape(B, cnt * Ls, K)
        if t1 == T - 1:
            Ez[:, -1] = ez[:, -1, Ls]
        if c0 == 0:
            Ez0 = e0.reshape(B, cnt, K)[:, 0].contiguous()
    if not batched:
        return logZ[0], (Ez0[0], Etrans[0], Ez[0])
    return logZ, (Ez0, Etrans, Ez)


class _HmmLogZParallelFn(torch.autograd.Function):
    """logZ with the reference's VJP, g * (E[z_0], E[trans], E[z_t]) (cython_hmm_inference.pyx:126-166), both
    computed chunk-parallel."""

    @staticmethod
    def forward(ctx, lp, lA, node, chunk_len):
        logZ, (Ez0, Etrans, Ez) = hmm_estep_parallel((lp, lA, node), chunk_len)
        ctx.save_for_backward(Ez0, Etrans, Ez)
        return logZ

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        Ez0, Etrans, Ez = ctx.saved_tensors
        return (g[:, None] * Ez0).sum(0), (g[:, None, None] * Etrans).sum(0), g[:, None, None] * Ez, None


def hmm_logZ_parallel(natparam, chunk_len=None):
    """hmm_logZ (svae/hmm/hmm_inference.py:12-18, 44-51) for few long sequences; differentiable in all three params."""
    lp, lA, node, batched, _ = H._canon(natparam)
    logZ = _HmmLogZParallelFn.apply(lp, lA, node, chunk_len)
    return logZ if batched else logZ[0]


def hmm_viterbi_parallel(natparam, chunk_len=None):
    """hmm_viterbi (svae/hmm/hmm_inference.py:54-63) for few long sequences: max-plus chunk matrices, back-tracked
    boundary states, then the ordinary Viterbi kernel on every chunk with both end states pinned.  The optimal path is
    unique for tie-free potentials, so the result equals the sequential one (tests/test_hmm_pit_gpu.py, exact).  With TIES
    the sequential kernel and this driver both return a maximum-score path but may break the tie differently (first
    argmax per step there; per chunk boundary, then per step here): equal scores, possibly different paths.  The -1e300
    pins rely on the Viterbi kernel adding node potentials without exponentiating or rescaling them (csrc/hmm_estep.cu)."""
    lp, lA, node, batched, dev = H._canon(natparam)
    B, T, K = node.shape
    Lc = int(chunk_len) if chunk_len else default_chunk_len(T)
    if T < 4 or Lc >= T - 1:
        return H.hmm_viterbi(natparam)
    L = _lib.lib()
    C = L.svae_lds_compose_num_chunks(T, Lc)
    V = torch.empty((B, C, K, K), dtype=torch.float64, device=dev)
    delta = torch.empty((B, C + 1, K), dtype=torch.float64, device=dev)
    zb = torch.empty((B, C + 1), dtype=torch.int32, device=dev)
    ptr, st = _lib.ptr, _lib.stream_ptr(dev)
    import ctypes
    _lib.check(L.svae_hmm_compose_chunks_maxplus(B, T, K, Lc, ptr(lA), ptr(node), ptr(V), st), "svae_hmm_compose_chunks_maxplus")
    last = node[:, -1].contiguous()
    _lib.check(L.svae_hmm_chain_viterbi(B, C, K, ptr(lp), ptr(V), ptr(last), ptr(delta), ctypes.c_void_p(zb.data_ptr()), st),
               "svae_hmm_chain_viterbi")
    never = -1e300
    pin = torch.full((B, C + 1, K), never, dtype=torch.float64, device=dev)
    pin.scatter_(2, zb.long()[..., None], 0.0)                       # 0 at the optimal boundary state, "never" elsewhere
    nfull = (T - 1) // Lc
    tail = (T - 1) - nfull * Lc
    path = torch.empty((B, T), dtype=torch.int32, device=dev)
    zero_pi = torch.zeros(K, dtype=torch.float64, device=dev)
    for c0, cnt, Ls in [(0, nfull, Lc)] + ([(nfull, 1, tail)] if tail else []):
        t0, t1 = c0 * Lc, c0 * Lc + cnt * Ls
        sub = torch.empty((B, cnt, Ls + 1, K), dtype=torch.float64, device=dev)
        sub[:, :, :Ls] = node[:, t0:t1].reshape(B, cnt, Ls, K)
        sub[:, :, 0] += pin[:, c0:c0 + cnt]
        sub[:, :, Ls] = pin[:, c0 + 1:c0 + cnt + 1]
        p = H.hmm_viterbi((zero_pi, lA, sub.reshape(B * cnt, Ls + 1, K))).reshape(B, cnt, Ls + 1)
        path[:, t0:t1] = p[:, :, :Ls].reshape(B, cnt * Ls)
        if t1 == T - 1:
            path[:, -1] = p[:, -1, Ls]
    return path if batched else path[0]
