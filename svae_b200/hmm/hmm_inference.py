"""Batched HMM local E-step on B200 -- host-side mirror of the reference's svae/hmm/hmm_inference.py.

Same names and return structure as the reference:
  hmm_logZ(natparam)      -> logZ            (hmm_inference.py:12-18; differentiable, VJP = E-step stats)
  hmm_estep(natparam)     -> (logZ, (E[z_0], E[transition counts], E[z_t]))       (hmm_inference.py:21-41)
  hmm_viterbi(natparam)   -> int32 state path                                      (hmm_inference.py:54-63)
  hmm_estep_slow          == hmm_estep       (hmm_inference.py:65, vgrad(hmm_logZ))
natparam = (init_params (K,), pair_params (K,K) row = from-state, node_params (T,K) or (B,T,K)), all log
values.  Everything runs in the sm_100a kernels of libsvae_b200.so; no CPU path.  Up to 16 states: the register-resident
kernels of csrc/hmm_estep.cu (lane = state); 17..32: 8 sequences per warp on the fp64 tensor-core instruction
(csrc/hmm_tc.cu); 33..1024: one CTA per sequence, thread = state (csrc/hmm_wide.cu) -- the reference takes any K.
"""
from __future__ import annotations
import ctypes
import torch
from .. import _lib


def _canon(natparam):
    init_params, pair_params, node_params = natparam
    dev = node_params.device if torch.is_tensor(node_params) and node_params.is_cuda else torch.device("cuda", torch.cuda.current_device())
    node = _lib.f64(node_params, dev)
    batched = node.dim() == 3
    if node.dim() not in (2, 3):
        raise ValueError("node_params must be (T,K) or (B,T,K)")
    if not batched:
        node = node[None]
    B, T, K = node.shape
    lp, lA = _lib.f64(init_params, dev), _lib.f64(pair_params, dev)
    if lp.shape != (K,) or lA.shape != (K, K):
        raise ValueError("init/pair params shapes %s %s for K=%d" % (tuple(lp.shape), tuple(lA.shape), K))
    if K > _lib.SVAE_HMM_WIDE_MAX_K:
        raise ValueError("K=%d > %d" % (K, _lib.SVAE_HMM_WIDE_MAX_K))
    return lp, lA, node.contiguous(), batched, dev


def hmm_estep_forward(lp, lA, node, logz_only=False, active=None, out=None):
    """One launch of the scaled forward(-backward) kernel on canonical (contiguous CUDA fp64) inputs.
    active: optional (B,) int32 mask -- sequences with 0 are skipped and keep what `out` = (logZ, Ez0, Etrans, Ez) held."""
    B, T, K = node.shape
    dev = node.device
    new = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)
    if out is not None:
        logZ, Ez0, Etrans, Ez = out
    else:
        logZ = new(B)
        Ez0 = Etrans = Ez = None
        if not logz_only:
            Ez0, Etrans, Ez = new(B, K), new(B, K, K), new(B, T, K)
    P = _lib.ptr
    L = _lib.lib()
    flags = _lib.SVAE_HMM_LOGZ_ONLY if logz_only else 0
    act = None if active is None else ctypes.c_void_p(active.data_ptr())
    if K >= _lib.SVAE_HMM_TC_MIN_K:
        # 17..32 states: tensor-core kernels (scratch: scaled forward messages and w/s, 2 B T K doubles);
        # more than 32: one CTA per sequence, thread = state (scratch: exp(log A - max) and its transpose)
        tc_logz = logz_only and K <= _lib.SVAE_HMM_MAX_K          # the tensor-core forward sweep alone stores nothing
        wbytes = 0 if tc_logz else L.svae_hmm_wide_workspace_bytes(B, T, K)
        work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
        code = L.svae_hmm_estep_wide(B, T, K, flags, P(lp), P(lA), P(node), act, P(logZ), P(Ez0), P(Etrans), P(Ez),
                                     ctypes.c_void_p(work.data_ptr()), wbytes, _lib.stream_ptr(dev))
        _lib.check(code, "svae_hmm_estep_wide")
        return logZ, Ez0, Etrans, Ez
    code = L.svae_hmm_estep_masked(B, T, K, flags, P(lp), P(lA), P(node), act,
                                   P(logZ), P(Ez0), P(Etrans), P(Ez), _lib.stream_ptr(dev))
    _lib.check(code, "svae_hmm_estep_masked")
    return logZ, Ez0, Etrans, Ez


class _HmmLogZFn(torch.autograd.Function):
    """logZ with the reference's VJP: g * (E[z_0], E[trans], E[z_t])  (cython_hmm_inference.pyx:126-166)."""

    @staticmethod
    def forward(ctx, lp, lA, node):
        logZ, Ez0, Etrans, Ez = hmm_estep_forward(lp, lA, node)
        ctx.save_for_backward(Ez0, Etrans, Ez)
        return logZ

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        Ez0, Etrans, Ez = ctx.saved_tensors
        return (g[:, None] * Ez0).sum(0), (g[:, None, None] * Etrans).sum(0), g[:, None, None] * Ez


def hmm_logZ(natparam):
    lp, lA, node, batched, _ = _canon(natparam)
    if torch.is_grad_enabled() and any(torch.is_tensor(x) and x.requires_grad for x in natparam):
        logZ = _HmmLogZFn.apply(lp, lA, node)
    else:
        logZ = hmm_estep_forward(lp, lA, node, logz_only=True)[0]
    return logZ if batched else logZ[0]


hmm_logZ_python = hmm_logZ


def hmm_estep(natparam):
    lp, lA, node, batched, _ = _canon(natparam)
    logZ, Ez0, Etrans, Ez = hmm_estep_forward(lp, lA, node)
    if not batched:
        return logZ[0], (Ez0[0], Etrans[0], Ez[0])
    return logZ, (Ez0, Etrans, Ez)


hmm_estep_slow = hmm_estep


def hmm_viterbi(natparam):
    lp, lA, node, batched, dev = _canon(natparam)
    B, T, K = node.shape
    L = _lib.lib()
    path = torch.empty((B, T), dtype=torch.int32, device=dev)
    wbytes = L.svae_hmm_viterbi_workspace_bytes(B, T, K)
    work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
    P = _lib.ptr
    code = L.svae_hmm_viterbi(B, T, K, P(lp), P(lA), P(node), ctypes.c_void_p(path.data_ptr()),
                              ctypes.c_void_p(work.data_ptr()), wbytes, _lib.stream_ptr(dev))
    _lib.check(code, "svae_hmm_viterbi")
    return path if batched else path[0]
