"""Parallel-in-time LDS E-step for the small-batch / long-T regime.

The fused kernel (lds_inference.py) walks a sequence step by step with one sub-warp: with B sequences only B tiles
of the GPU work, however long T is.  This driver cuts every sequence into chunks and gets the SAME numbers --
statistics, log-normaliser and, for the same noise, the same samples as natural_lds_inference_general
(svae/lds/lds_inference.py:185-202) -- from three wide launches instead of one narrow one:

  1. svae_lds_compose_chunks: every chunk's node and pair potentials integrated into one pair potential on the
     chunk's end states (B x C tiles);
  2. the fused kernel on the chain of chunk boundaries (B sequences of C+1 states, inhomogeneous pair parameters):
     total log-normaliser, prediction messages and smoothed moments at every boundary;
  3. the fused kernel on all chunks at once (B x C sequences of chunk_len+1 states): init = the chain's prediction
     message at the chunk start (SVAE_LDS_INIT_PER_SEQ), last node potential = smoothed / predicted at the chunk end
     (term_J port), so every chunk's smoother returns the exact posterior moments of its steps.
     Sample paths: a chunk's backward sampler is affine in its terminal draw, x_t = x_t^k + G_{t->e} (x_e - x_e^k);
     n+1 probe lanes (zero noise, unit terminal noise) measure G_{t->e} L_e^-T, and a log-depth affine suffix scan
     over the chunks replaces every terminal draw by the next chunk's first state -- which reproduces the sequential
     recursion x_t = b_t + G_t x_{t+1} (cython_lds_inference.pyx:335-352) exactly.

Forward values only (no VJP); diagonal node potentials; pair parameters homogeneous, per-time or per-sequence (the
SLDS case, svae/models/slds_svae.py:92-103).  The torch calls below are
reshapes, two batched n x n inverses at the C boundaries and the O(log C) stitch -- the arithmetic is in the kernels.
"""
from __future__ import annotations
import ctypes
import math
import torch
from .. import _lib
from . import lds_inference as K


def _dev(h_any):
    return h_any.device if torch.is_tensor(h_any) and h_any.is_cuda else torch.device("cuda", torch.cuda.current_device())


def compose_chunks(pair_params, node_params, chunk_len):
    """Stage 1 alone.  Returns ((cJ11, cJ12, cJ22, clogZ) natural form (B,C,...), e1, e2 (B,C,n))."""
    device = _dev(node_params[1])
    nJ, nh, nlogZ, _ = K._canonical_node_params(node_params, device)
    if nJ.dim() != 3:
        raise ValueError("parallel-in-time driver: diagonal node potentials only")
    B, T, n = nh.shape
    J11, J12, J22, plogZ, st, sb = K._canonical_pair_params(pair_params, B, T, n, device)
    L = _lib.lib()
    C = L.svae_lds_compose_num_chunks(T, int(chunk_len))
    new = lambda *s: torch.empty(s, dtype=torch.float64, device=device)
    cJ11, cJ12, cJ22, clogZ, e1, e2 = new(B, C, n, n), new(B, C, n, n), new(B, C, n, n), new(B, C), new(B, C, n), new(B, C, n)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    P = _lib.ptr
    code = L.svae_lds_compose_chunks(B, T, n, int(chunk_len), P(J11), P(J12), P(J22), P(plogZ), st, sb, P(nJ), P(nh), P(nlogZ),
                                     P(cJ11), P(cJ12), P(cJ22), P(clogZ), P(e1), P(e2),
                                     ctypes.c_void_p(status.data_ptr()), _lib.stream_ptr(device))
    _lib.check(code, "svae_lds_compose_chunks")
    compose_chunks.last_status = status          # device word: 1 + first sequence with a non-positive pivot
    return (cJ11, cJ12, cJ22, clogZ), e1, e2


def default_chunk_len(T):
    """~sqrt(T-1), nudged to a divisor of T-1 when one is within a factor 1.5: no ragged tail chunk means one launch
    group instead of two (the driver's cost at moderate T is its launch count)."""
    target = max(2, int(round(math.sqrt(max(T - 1, 1)))))
    best = None
    for L in range(max(2, int(target / 1.5)), int(target * 1.5) + 1):
        if (T - 1) % L == 0 and (best is None or abs(L - target) < abs(best - target)):
            best = L
    return best or target


def _affine_suffix_scan(A, d):
    """delta_c = d_c + A_c delta_{c+1}, delta beyond the end = 0.  A (B,C,n,n), d (B,C,S,n) -> delta (B,C,S,n).
    Hillis-Steele doubling over the chunk axis: ceil(log2 C) rounds of batched products."""
    C = A.shape[1]
    k = 1
    pad = torch.nn.functional.pad
    while k < C:
        d_sh = pad(d[:, k:], (0, 0, 0, 0, 0, k))           # shift along the chunk axis, zeros beyond the end
        A_sh = pad(A[:, k:], (0, 0, 0, 0, 0, k))
        d = torch.baddbmm(d.reshape(-1, d.shape[2], d.shape[3]), d_sh.reshape(-1, d.shape[2], d.shape[3]),
                          A.reshape(-1, A.shape[2], A.shape[3]).transpose(1, 2)).reshape(d.shape)
        A = A @ A_sh
        k *= 2
    return d


def natural_lds_inference_parallel(natparam, node_params, num_samples=1, eps=None, chunk_len=None, generator=None,
                                   check_status=False):
    """(samples, (E_init_stats, E_pair_stats, E_node_stats), lognorm): the contract of natural_lds_inference_general
    (svae/lds/lds_inference.py:185-202), computed chunk-parallel.  chunk_len defaults to ~sqrt(T).
    check_status: synchronise and raise FloatingPointError if any stage met a non-positive-definite pivot (the
    reference's asserts, svae/lds/gaussian.py:48,64)."""
    init_params, pair_params = natparam
    if any(torch.is_tensor(x) and x.requires_grad for x in node_params) and torch.is_grad_enabled():
        raise RuntimeError("natural_lds_inference_parallel is forward-only; use natural_lds_inference_general for gradients")
    device = _dev(node_params[1])
    nJ, nh, nlogZ, batched = K._canonical_node_params(node_params, device)
    if nJ.dim() != 3:
        raise ValueError("parallel-in-time driver: diagonal node potentials only")
    B, T, n = nh.shape
    S = int(num_samples) if num_samples else 0
    Lc = int(chunk_len) if chunk_len else default_chunk_len(T)
    if T < 4 or Lc >= T - 1:
        return K.natural_lds_inference_general(natparam, node_params, num_samples, eps, generator)
    if S > 0:
        if eps is None:
            eps = torch.randn((B, T, S, n), dtype=torch.float64, device=device, generator=generator)
        else:
            eps = _lib.f64(eps, device)
            eps = eps[None] if (not batched and eps.dim() == 3) else eps
            if eps.shape != (B, T, S, n):
                raise ValueError("eps must be (B,T,S,n)")
    homog = K._is_homog(pair_params)
    if not homog:                            # per-time / per-sequence pair parameters (the SLDS case), as (B,T-1,...)
        pJ11, pJ12, pJ22, plogZ, _, _ = K._canonical_pair_params(pair_params, B, T, n, device)
        pfull = tuple(x.expand(B, T - 1, n, n) for x in (pJ11, pJ12, pJ22))
        plz = None if plogZ is None else plogZ.reshape((-1, T - 1)).expand(B, T - 1)
    lanes = 32 - (n + 1)                     # sample lanes one tile can carry next to the n+1 filter lanes
    if S > lanes:
        raise ValueError("num_samples %d > %d" % (S, lanes))
    f64 = dict(dtype=torch.float64, device=device)

    # ---- 1. chunk elements -------------------------------------------------------------------------------
    node = (nJ, nh) if nlogZ is None else (nJ, nh, nlogZ)
    cpair, e1, e2 = compose_chunks(pair_params, node, Lc)
    C = e1.shape[1]
    nfull = (T - 1) // Lc                    # chunks of exactly Lc steps; a shorter tail chunk if C > nfull
    tail = (T - 1) - nfull * Lc

    # ---- 2. boundary chain -------------------------------------------------------------------------------
    ch = torch.zeros((B, C + 1, n), **f64)
    ch[:, :-1] += e1
    ch[:, 1:] += e2
    ch[:, -1] += nh[:, -1]
    cJ = torch.zeros((B, C + 1, n), **f64)
    cJ[:, -1] = nJ[:, -1]
    clZ = torch.zeros((B, C + 1), **f64)
    if nlogZ is not None:
        clZ[:, -1] = nlogZ[:, -1]
    chain = K.lds_estep_forward(init_params, cpair, (cJ, ch, clZ), 0, None, per_t_pair_stats="xx", want_stats=False,
                                messages_out=True)
    lognorm = chain["lognorm"]
    Jp, hp = chain["J_pred"], chain["h_pred"].clone()
    hp[:, 1:] += e2                          # eta2 of chunk c-1 rode as node potential of boundary c
    mu = chain["E_node_x"]
    Sig = chain["Exx_t"] - mu[..., :, None] * mu[..., None, :]
    Js = torch.linalg.inv_ex(Sig[:, 1:])[0]        # _ex: no host sync, no exception on a failed sequence
    Js = 0.5 * (Js + Js.transpose(-1, -2))
    tJ = -0.5 * Js - Jp[:, 1:]               # terminal node potential of chunk c: smoothed / predicted at its end
    th = torch.einsum('bcij,bcj->bci', Js, mu[:, 1:]) - hp[:, 1:]

    # ---- 3. all chunks as independent sequences ----------------------------------------------------------------
    groups = [(0, nfull, Lc)] + ([(nfull, 1, tail)] if tail else [])     # (first chunk, count, length)
    E_node_xx, E_node_x = torch.empty((B, T, n), **f64), torch.empty((B, T, n), **f64)
    samples = torch.empty((B, T, S, n), **f64) if S else None
    E_pair = [torch.zeros((B, n, n), **f64) for _ in range(3)]
    Exx_t = None if homog else torch.empty((B, T, n, n), **f64)
    Exxn_t = None if homog else torch.empty((B, T - 1, n, n), **f64)
    E_init = None
    x_start = torch.empty((B, C, S, n), **f64) if S else None       # kernel paths at the chunk ends, and G_{s->e}
    x_end = torch.empty((B, C, S, n), **f64) if S else None
    R_start = torch.empty((B, C, n, n), **f64) if S else None
    per_group, statuses = [], []
    for c0, cnt, L in groups:
        Q = B * cnt
        t0, t1 = c0 * Lc, c0 * Lc + cnt * L
        sJ = torch.zeros((B, cnt, L + 1, n), **f64)
        sh = torch.zeros((B, cnt, L + 1, n), **f64)
        sJ[:, :, :L] = nJ[:, t0:t1].reshape(B, cnt, L, n)
        sh[:, :, :L] = nh[:, t0:t1].reshape(B, cnt, L, n)
        sh[:, :, L] = th[:, c0:c0 + cnt]
        sub_init = (Jp[:, c0:c0 + cnt].reshape(Q, n, n), hp[:, c0:c0 + cnt].reshape(Q, n))
        sub_node = (sJ.reshape(Q, L + 1, n), sh.reshape(Q, L + 1, n))
        sub_term = tJ[:, c0:c0 + cnt].reshape(Q, n, n)
        if homog:
            sub_pair = pair_params
        else:
            sub_pair = tuple(x[:, t0:t1].reshape(Q, L, n, n) for x in pfull)
            sub_pair += (torch.zeros((), **f64) if plz is None else plz[:, t0:t1].reshape(Q, L),)

        def window(x):                       # (B,T,...) -> (B,cnt,L+1,...): steps t0 + c L .. t0 + (c+1) L inclusive
            body = x[:, t0:t1].reshape((B, cnt, L) + x.shape[2:])
            last = x[:, t0 + L:t1 + 1:L].reshape((B, cnt, 1) + x.shape[2:])
            return torch.cat([body, last], 2)

        probes = 0
        if S:
            probes = min(n + 1, lanes - S)   # probe lanes riding in the statistics launch (0 = separate launches)
            if probes < 2:
                probes = 0
            SP = S + probes
            seps = torch.zeros((B, cnt, L + 1, SP, n), **f64)
            seps[:, :, :, :S] = window(eps)
            if probes:
                seps[:, :, L, S + 1:] = torch.eye(n, **f64)[:probes - 1]
            o = K.lds_estep_forward(sub_init, sub_pair, sub_node, SP, seps.reshape(Q, L + 1, SP, n), term_J=sub_term,
                                    per_t_pair_stats=not homog)
            xs = o["samples"].reshape(B, cnt, L + 1, SP, n)
            xk = xs[..., :S, :]
            cols = [xs[..., S + 1:, :] - xs[..., S:S + 1, :]] if probes else []
            done = max(probes - 1, 0)
            while done < n:                  # remaining unit probes: sampling-only launches
                k = min(n - done, lanes - 1)
                peps = torch.zeros((Q, L + 1, k + 1, n), **f64)
                peps[:, L, 1:] = torch.eye(n, **f64)[done:done + k]
                po = K.lds_estep_forward(sub_init, sub_pair, sub_node, k + 1, peps, want_stats=False, term_J=sub_term)
                ps = po["samples"].reshape(B, cnt, L + 1, k + 1, n)
                cols.append(ps[..., 1:, :] - ps[..., :1, :])
                done += k
            D = torch.cat(cols, -2).transpose(-1, -2)                      # (B,cnt,L+1,n,n): column j = probe j
            Rt = D @ torch.linalg.inv_ex(D[:, :, L])[0][:, :, None]              # G_{t->e}
            x_start[:, c0:c0 + cnt], x_end[:, c0:c0 + cnt] = xk[:, :, 0], xk[:, :, L]
            R_start[:, c0:c0 + cnt] = Rt[:, :, 0]
        else:
            o = K.lds_estep_forward(sub_init, sub_pair, sub_node, 0, None, term_J=sub_term, per_t_pair_stats=not homog)
            xk = Rt = None
        statuses.append(o["status"])
        if homog:
            for k, key in enumerate(("E_pair_xx", "E_pair_xxn", "E_pair_xnxn")):
                E_pair[k] += o[key].reshape(B, cnt, n, n).sum(1)
        else:
            xx = o["Exx_t"].reshape(B, cnt, L + 1, n, n)
            Exx_t[:, t0:t1] = xx[:, :, :L].reshape(B, cnt * L, n, n)
            if t1 == T - 1:
                Exx_t[:, -1] = xx[:, -1, L]
            Exxn_t[:, t0:t1] = o["Exxn_t"].reshape(B, cnt * L, n, n)
        if c0 == 0:
            E_init = (o["E_init_xx"].reshape(B, cnt, n, n)[:, 0].contiguous(), o["E_init_x"].reshape(B, cnt, n)[:, 0].contiguous())
        exx, ex = o["E_node_xx"].reshape(B, cnt, L + 1, n), o["E_node_x"].reshape(B, cnt, L + 1, n)
        E_node_xx[:, t0:t1], E_node_x[:, t0:t1] = exx[:, :, :L].reshape(B, cnt * L, n), ex[:, :, :L].reshape(B, cnt * L, n)
        if t1 == T - 1:
            E_node_xx[:, -1], E_node_x[:, -1] = exx[:, -1, L], ex[:, -1, L]
        per_group.append((c0, cnt, L, t0, t1, xk, Rt))

    if S:
        # delta_c = (true first state of chunk c+1) - (terminal draw of chunk c); zero for the last chunk
        A = torch.zeros((B, C, n, n), **f64)
        d = torch.zeros((B, C, S, n), **f64)
        A[:, :-1] = R_start[:, 1:]
        d[:, :-1] = x_start[:, 1:] - x_end[:, :-1]
        delta = _affine_suffix_scan(A, d)
        for c0, cnt, L, t0, t1, xk, Rt in per_group:
            fixed = xk + torch.einsum('bctij,bcsj->bctsi', Rt, delta[:, c0:c0 + cnt])
            samples[:, t0:t1] = fixed[:, :, :L].reshape(B, cnt * L, S, n)
            if t1 == T - 1:
                samples[:, -1] = fixed[:, -1, L]

    if check_status:
        bad = int(compose_chunks.last_status.item()) or int(chain["status"].item()) or max(int(x.item()) for x in statuses)
        if bad:
            raise FloatingPointError("parallel-in-time LDS E-step: non-positive-definite pivot (stage-local sequence %d)" % (bad - 1))
    o = dict(B=B, T=T, n=n, S=S, batched=batched, lognorm=lognorm, samples=samples, E_node_xx=E_node_xx, E_node_x=E_node_x,
             E_init_xx=E_init[0], E_init_x=E_init[1], E_pair_xx=E_pair[0], E_pair_xxn=E_pair[1], E_pair_xnxn=E_pair[2],
             Exx_t=Exx_t, Exxn_t=Exxn_t)
    return K._pack_stats(o, homog)
