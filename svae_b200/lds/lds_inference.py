"""Batched LDS local E-step on B200 -- host-side mirror of the reference's svae/lds/lds_inference.py.

Same names, argument meaning and return structure as the reference drivers
  natural_lds_inference_general / cython_natural_lds_inference_general   (lds_inference.py:185-202)
  natural_lds_estep_general / cython_natural_lds_estep_general           (lds_inference.py:223-237)
  natural_lds_sample / cython_natural_lds_sample                         (lds_inference.py:253-264)
with one addition: node parameters may carry a leading batch axis of B independent sequences
(the reference handles one sequence per call).  Unbatched (T,n) inputs are accepted and return
unbatched results exactly like the reference.

All arithmetic runs in the hand-written sm_100a kernels of libsvae_b200.so through the C-ABI in
include/svae_b200.h; PyTorch only owns memory and streams.  No CPU path exists.
"""
from __future__ import annotations
import ctypes
import torch
from .. import _lib


# ------------------------------------------------------------------------------------------------
# parameter bookkeeping (lds_inference.py:44-80)
# ------------------------------------------------------------------------------------------------

def _canonical_init_params(init_params, device):
    """(J, h, logZ...) -> (J, h, sum of the scalar terms) as in lds_inference.py:61-62."""
    J = _lib.f64(init_params[0], device)
    h = _lib.f64(init_params[1], device)
    logZ = None
    for term in init_params[2:]:
        term = _lib.f64(term, device)
        logZ = term if logZ is None else logZ + term
    return J, h, (None if logZ is None else logZ.reshape(-1).contiguous())


def _canonical_node_params(node_params, device):
    """Accepts (J,h) or (J,h,logZ); h (T,n) or (B,T,n); J diagonal (same shape as h) or dense (...,T,n,n)
    (lds_inference.py:65-80: ndims (3,2,1),(3,2),(2,2,1),(2,2)).  The reference's Cython path is diag-only
    (cython_lds_inference.pyx:43); dense potentials follow its pure-Python path (lds/gaussian.py:46-49) and
    make E_node_stats[0] the full E[x x'] (lds_inference.py:159-167).  Other shapes: ValueError (:80).
    """
    if not isinstance(node_params, (tuple, list)) or len(node_params) not in (2, 3):
        raise ValueError("node_params must be (J, h) or (J, h, logZ)")
    J = _lib.f64(node_params[0], device)
    h = _lib.f64(node_params[1], device)
    batched = h.dim() == 3
    if h.dim() not in (2, 3) or (J.shape != h.shape and J.shape != h.shape + h.shape[-1:]):
        raise ValueError("node_params shapes %s %s not supported (h (T,n) or (B,T,n); J like h or dense)"
                         % (tuple(J.shape), tuple(h.shape)))
    if not batched:
        J, h = J[None], h[None]
    logZ = None
    if len(node_params) == 3 and node_params[2] is not None:
        logZ = _lib.f64(node_params[2], device)
        if not batched:
            logZ = logZ[None]
        if logZ.shape != h.shape[:2]:
            raise ValueError("node logZ shape %s" % (tuple(logZ.shape),))
    return J.contiguous(), h.contiguous(), logZ, batched


def _canonical_pair_params(pair_params, B, T, n, device):
    """Homogeneous (n,n) blocks, time-inhomogeneous (T-1,n,n) or per-sequence (B,T-1,n,n)
    (_repeat_param, lds_inference.py:50-59; ensure_time_axis, cython_lds_inference.pyx:17-26)."""
    J11, J12, J22 = (_lib.f64(x, device) for x in pair_params[:3])
    logZ = _lib.f64(pair_params[3], device) if len(pair_params) > 3 and pair_params[3] is not None else None
    if not (J11.shape == J12.shape == J22.shape) or J11.shape[-2:] != (n, n):
        raise ValueError("pair_params shapes")
    if J11.dim() == 2:
        st, sb = 0, 0
        if logZ is not None:
            logZ = logZ.reshape(1)
    elif J11.dim() == 3 and J11.shape[0] == T - 1:
        st, sb = n * n, 0
        if logZ is not None:
            logZ = logZ.expand(T - 1).contiguous() if logZ.dim() == 0 else logZ.reshape(T - 1)
    elif J11.dim() == 4 and J11.shape[:2] == (B, T - 1):
        st, sb = n * n, (T - 1) * n * n
        if logZ is not None:
            logZ = logZ.expand(B, T - 1).contiguous()
    else:
        raise ValueError("pair_params leading shape %s" % (tuple(J11.shape[:-2]),))
    return J11, J12, J22, logZ, st, sb


# ------------------------------------------------------------------------------------------------
# the kernel call
# ------------------------------------------------------------------------------------------------

def lds_estep_forward(init_params, pair_params, node_params, num_samples=1, eps=None,
                      per_t_pair_stats=False, want_stats=True, generator=None, check_status=False,
                      save_for_bwd=False, buffers=None, messages_out=False, from_messages=None, term_J=None, active=None):
    """One launch of the fused filter + smoother + sampler kernel.  Returns a dict of raw tensors.
    With save_for_bwd the workspace keeps what lds_estep_backward needs (returned as out["workspace"]).
    messages_out: also return the prediction messages (out["J_pred"], out["h_pred"]).
    from_messages = (J_filt (B,T,n,n), h_filt (B,T,n)): run the backward sweeps from these filtered messages
    instead of the kernel's own (node_params must then be zero).
    term_J (B,n,n): dense natural-form precision added to the last step's node potential (lds_pit.py).
    active (B,) int32: sequences with 0 are skipped; their rows of every output keep what `buffers` held."""
    h_any = node_params[1]
    device = h_any.device if torch.is_tensor(h_any) and h_any.is_cuda else torch.device("cuda", torch.cuda.current_device())
    nJ, nh, nlogZ, batched = _canonical_node_params(node_params, device)
    B, T, n = nh.shape
    dense, nJd = nJ.dim() == 4, None
    if dense:                      # dense precisions ride the cold port; the diagonal port gets zeros
        nJd, nJ = nJ, torch.zeros_like(nh)
    if n > _lib.SVAE_LDS_MAX_N:
        raise ValueError("latent dimension %d > %d: the fused kernel is compiled for n <= %d; the reference-named drivers "
                         "and primitives (natural_lds_inference_general, natural_lds_estep_general, natural_lds_sample, "
                         "natural_filter_forward_general, natural_smoother_general, natural_sample_backward) take larger n "
                         "through lds_large.py" % (n, _lib.SVAE_LDS_MAX_N, _lib.SVAE_LDS_MAX_N))
    iJ, ih, ilogZ = _canonical_init_params(init_params, device)
    init_per_seq = iJ.dim() == 3
    if init_per_seq:
        if iJ.shape != (B, n, n) or ih.shape != (B, n) or (ilogZ is not None and ilogZ.numel() != B):
            raise ValueError("per-sequence init_params shapes")
    elif iJ.shape != (n, n) or ih.shape != (n,):
        raise ValueError("init_params shapes")
    J11, J12, J22, plogZ, st, sb = _canonical_pair_params(pair_params, B, T, n, device)
    S = int(num_samples) if num_samples else 0
    if eps is not None:
        eps = _lib.f64(eps, device)
        if not batched and eps.dim() == 3:
            eps = eps[None]
        if eps.shape != (B, T, S, n):
            raise ValueError("eps must be (B,T,S,n) = %s, got %s" % ((B, T, S, n), tuple(eps.shape)))
        eps = eps.contiguous()
    elif S > 0:
        eps = torch.randn((B, T, S, n), dtype=torch.float64, device=device, generator=generator)

    L = _lib.lib()
    flags = (_lib.SVAE_LDS_PER_T_PAIR_STATS if (per_t_pair_stats or dense) else 0) | (_lib.SVAE_LDS_SAVE_FOR_BWD if save_for_bwd else 0) \
        | (_lib.SVAE_LDS_INIT_PER_SEQ if init_per_seq else 0)
    _pool = dict(buffers) if buffers else {}

    def new(*shape, name=None):
        t = _pool.get(name) if name else None
        if t is not None and tuple(t.shape) == tuple(shape):
            return t
        return torch.empty(shape, dtype=torch.float64, device=device)
    out = dict(B=B, T=T, n=n, S=S, batched=batched, eps=eps)
    out["lognorm"] = new(B, name="lognorm")
    out["E_node_xx"], out["E_node_x"] = new(B, T, n, name="E_node_xx"), new(B, T, n, name="E_node_x")
    out["samples"] = new(B, T, S, n, name="samples") if S > 0 else None
    if want_stats:
        out["E_init_xx"], out["E_init_x"] = new(B, n, n, name="E_init_xx"), new(B, n, name="E_init_x")
        out["E_pair_xx"], out["E_pair_xxn"], out["E_pair_xnxn"] = \
            new(B, n, n, name="E_pair_xx"), new(B, n, n, name="E_pair_xxn"), new(B, n, n, name="E_pair_xnxn")
    if per_t_pair_stats or dense:
        out["Exx_t"] = new(B, T, n, n, name="Exx_t")
    if per_t_pair_stats and per_t_pair_stats != "xx":     # "xx": the per-time second moments alone (boundary chain)
        out["Exxn_t"] = new(B, max(T - 1, 0), n, n, name="Exxn_t")
    if messages_out:
        out["J_pred"], out["h_pred"] = new(B, T, n, n), new(B, T, n)
    inJ = inh = None
    if from_messages is not None:
        inJ, inh = (_lib.f64(x, device).contiguous() for x in from_messages)
        if inJ.shape != (B, T, n, n) or inh.shape != (B, T, n):
            raise ValueError("from_messages must be ((B,T,n,n), (B,T,n))")
    if term_J is not None:
        term_J = _lib.f64(term_J, device).contiguous()
        if term_J.shape != (B, n, n):
            raise ValueError("term_J must be (B,n,n)")
    status = _pool.get("status")
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=device)
    wbytes = L.svae_lds_estep_workspace_bytes(B, T, n, S, flags)
    work = _pool.get("workspace")
    if work is None or work.numel() < wbytes:
        work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=device)
    g = out.get
    P = _lib.ptr
    code = L.svae_lds_estep_fwd_masked(
        B, T, n, S, flags,
        P(iJ), P(ih), P(ilogZ),
        P(J11), P(J12), P(J22), P(plogZ), st, sb,
        P(nJ), P(nh), P(nlogZ), P(eps),
        P(g("E_init_xx")), P(g("E_init_x")),
        P(g("E_pair_xx")), P(g("E_pair_xxn")), P(g("E_pair_xnxn")),
        P(out["E_node_xx"]), P(out["E_node_x"]),
        P(out["samples"]), P(out["lognorm"]),
        P(g("Exx_t")), P(g("Exxn_t")),
        P(g("J_pred")), P(g("h_pred")), P(inJ), P(inh), P(nJd), P(term_J),
        None if active is None else ctypes.c_void_p(active.data_ptr()),
        ctypes.c_void_p(status.data_ptr()), ctypes.c_void_p(work.data_ptr()), wbytes,
        _lib.stream_ptr(device))
    _lib.check(code, "svae_lds_estep_fwd_masked")
    out["status"] = status
    out["dense"] = dense
    if dense:
        out["E_node_xx"] = out["Exx_t"]      # full second moments (lds_inference.py:164-167)
    if save_for_bwd:
        out["workspace"], out["workspace_bytes"] = work, wbytes
    if check_status:
        bad = int(status.item())
        if bad:
            raise FloatingPointError("LDS E-step: non-positive-definite pivot in sequence %d "
                                     "(reference asserts: svae/lds/gaussian.py:48,64)" % (bad - 1))
    return out


def lds_reduce_global_stats(o, out=None, work=None):
    """sum over the minibatch of [E_init_xx | E_init_x | E_pair_xx | E_pair_xxn | E_pair_xnxn | lognorm]
    (flat(stats) of svae.py:33-34 plus the normaliser) with the deterministic two-stage CUDA reduction."""
    B, n = o["B"], o["n"]
    dev = o["lognorm"].device
    L = _lib.lib()
    if out is None:
        out = torch.empty(4 * n * n + n + 1, dtype=torch.float64, device=dev)
    wbytes = L.svae_lds_reduce_stats_workspace_bytes(n)
    if work is None:
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    P = _lib.ptr
    code = L.svae_lds_reduce_stats(B, n, P(o["E_init_xx"]), P(o["E_init_x"]), P(o["E_pair_xx"]), P(o["E_pair_xxn"]),
                                   P(o["E_pair_xnxn"]), P(o["lognorm"]), P(out), ctypes.c_void_p(work.data_ptr()),
                                   wbytes, _lib.stream_ptr(dev))
    _lib.check(code, "svae_lds_reduce_stats")
    return out


class LdsHostPipeline:
    """End-to-end E-step for HOST-resident minibatches: the public call a data loader would make.

    All device buffers (inputs, noise, outputs, stash workspace) and the streams are allocated ONCE for a
    given (B, T, n, S); each call cuts the batch into `chunks` pieces that flow
        pinned host -> H2D copy -> fused E-step kernel -> statistics reduction -> D2H copy -> pinned host
    on rotating streams, so PCIe transfers in both directions overlap the kernels of neighbouring chunks.
    """

    def __init__(self, B, T, n, S, chunks=3, device=None):
        """chunks: an int (equal pieces) or a sequence of piece sizes summing to B.  The pipeline is bound by the
        device-to-host copy (48 KB per sequence at n=10, T=200 against 32 KB in), which cannot start before the first
        piece's kernel has finished: a small first piece and growing later ones (ramp_schedule) start it earlier."""
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        if isinstance(chunks, int):
            sizes = [(B * (k + 1)) // chunks - (B * k) // chunks for k in range(chunks)]
        else:
            sizes = [int(x) for x in chunks]
            if sum(sizes) != B or min(sizes) < 0:
                raise ValueError("chunk sizes must be non-negative and sum to B")
        self.B, self.T, self.n, self.S, self.chunks = B, T, n, S, len(sizes)
        self.bounds = [0]
        for x in sizes:
            self.bounds.append(self.bounds[-1] + x)
        mk = lambda: torch.cuda.Stream(device=self.device)
        self.s_in, self.s_out = mk(), mk()                 # one copy stream per direction, kernels on rotating streams
        self.streams = [mk() for _ in range(min(3, len(sizes)))]
        self.ev_in = [torch.cuda.Event() for _ in sizes]
        self.ev_done = [torch.cuda.Event() for _ in sizes]
        new = lambda *s: torch.empty(s, dtype=torch.float64, device=self.device)
        L = _lib.lib()
        self.bufs = []
        for b in sizes:
            bb = max(b, 1)
            d = dict(J=new(bb, T, n), h=new(bb, T, n), eps=new(bb, T, S, n) if S else None,
                     lognorm=new(bb), E_node_xx=new(bb, T, n), E_node_x=new(bb, T, n),
                     samples=new(bb, T, S, n) if S else None, E_init_xx=new(bb, n, n), E_init_x=new(bb, n),
                     E_pair_xx=new(bb, n, n), E_pair_xxn=new(bb, n, n), E_pair_xnxn=new(bb, n, n),
                     status=torch.zeros(1, dtype=torch.int32, device=self.device),
                     workspace=torch.empty(max(L.svae_lds_estep_workspace_bytes(bb, T, n, S, 0), 16), dtype=torch.uint8,
                                           device=self.device),
                     acc=new(4 * n * n + n + 1),
                     rwork=torch.empty(L.svae_lds_reduce_stats_workspace_bytes(n), dtype=torch.uint8, device=self.device))
            self.bufs.append(d)
        self.acc = new(4 * n * n + n + 1)

    def run(self, natparam, node_host, out_host, generator=None, outputs="all"):
        """node_host = (J_node, h_node) pinned CPU tensors (B,T,n); out_host = dict of pinned CPU tensors filled in place.
        outputs="all": {"samples", "E_node_xx", "E_node_x", "lognorm"} -- every per-sequence output of
            natural_lds_inference_general (lds_inference.py:196-202) goes back to the host;
        outputs="run_inference": {"samples"} only -- what models/lds.py:35-41 run_inference hands its caller: the samples, the
            summed global statistics and local_kl = <node potentials, E_node> - lognorm; the node statistics are consumed on
            the device by that contraction (their only consumer) and never cross PCIe.
        Returns the device tensor sum_b [E_init | E_pair | lognorm] (see lds_reduce_global_stats); in "run_inference" mode
        (acc, local_kl) with local_kl a device scalar."""
        init_params, pair_params = natparam
        ri = outputs == "run_inference"
        J_h, h_h = node_host
        cur = torch.cuda.current_stream(self.device)
        S = self.S
        self.s_in.wait_stream(cur)
        self.s_out.wait_stream(cur)
        for st in self.streams:
            st.wait_stream(cur)
        live = [k for k in range(self.chunks) if self.bounds[k + 1] > self.bounds[k]]
        for k in live:
            d, lo, hi = self.bufs[k], self.bounds[k], self.bounds[k + 1]
            with torch.cuda.stream(self.s_in):
                d["J"].copy_(J_h[lo:hi], non_blocking=True)
                d["h"].copy_(h_h[lo:hi], non_blocking=True)
                self.ev_in[k].record(self.s_in)
            st = self.streams[k % len(self.streams)]
            st.wait_event(self.ev_in[k])
            with torch.cuda.stream(st):
                if S:
                    d["eps"].normal_(generator=generator)
                o = lds_estep_forward(init_params, pair_params, (d["J"], d["h"]), S, d["eps"], buffers=d)
                lds_reduce_global_stats(o, out=d["acc"], work=d["rwork"])
                if ri:      # <node potentials, E_node> of this piece (models/lds.py:40); lognorm's sum is the last entry of acc
                    d["kl"] = torch.dot(d["J"].reshape(-1), o["E_node_xx"].reshape(-1)) + torch.dot(d["h"].reshape(-1), o["E_node_x"].reshape(-1))
                self.ev_done[k].record(st)
            self.s_out.wait_event(self.ev_done[k])
            with torch.cuda.stream(self.s_out):
                if S:
                    out_host["samples"][lo:hi].copy_(o["samples"], non_blocking=True)
                if not ri:
                    out_host["E_node_xx"][lo:hi].copy_(o["E_node_xx"], non_blocking=True)
                    out_host["E_node_x"][lo:hi].copy_(o["E_node_x"], non_blocking=True)
                    out_host["lognorm"][lo:hi].copy_(o["lognorm"], non_blocking=True)
        self.acc.zero_()
        cur.wait_stream(self.s_out)
        for st in self.streams:
            cur.wait_stream(st)
        for k in live:
            self.acc += self.bufs[k]["acc"]
        if ri:
            kl = self.bufs[live[0]]["kl"]
            for k in live[1:]:
                kl = kl + self.bufs[k]["kl"]
            return self.acc, kl - self.acc[-1]
        return self.acc


def ramp_schedule(B, first=256, growth=1.5, largest=1024):
    """Piece sizes first, first*growth, ... capped at `largest`, summing to B (the last piece takes the remainder)."""
    sizes, x, left = [], float(first), B
    while left > 0:
        b = min(left, int(min(x, largest)))
        if left - b < first // 2:      # do not leave a sliver
            b = left
        sizes.append(b)
        left -= b
        x *= growth
    return sizes


def trapezoid_schedule(B):
    """Piece sizes for the host pipeline when the host-to-device copy is the bound (run_inference mode): small first piece
    (the first kernel starts early), plateau, small last pieces (after the last input byte has arrived only a short kernel and
    a short copy remain).  Swept on a B200 (scripts/e2e_probe_ri.py, profiles/r02_e2e_chunk_sweep.txt): 3.47 ms against
    3.84 ms for the ramp at B = 4096."""
    if B < 1024:
        return [B]
    fr = (0.09375, 0.1875, 0.25, 0.25, 0.15625)
    sizes = [max(1, int(round(f * B))) for f in fr]
    sizes.append(B - sum(sizes))
    return [x for x in sizes if x > 0]


def lds_estep_host(natparam, node_host, num_samples, out_host, chunks=3, generator=None, device=None, _plans={}, outputs="all"):
    """Functional wrapper around LdsHostPipeline (plans are cached per shape)."""
    B, T, n = node_host[1].shape
    S = int(num_samples) if num_samples else 0
    device = device or torch.device("cuda", torch.cuda.current_device())
    key = (B, T, n, S, chunks if isinstance(chunks, int) else tuple(chunks), str(device))
    if key not in _plans:
        _plans[key] = LdsHostPipeline(B, T, n, S, chunks, device)
    return _plans[key].run(natparam, node_host, out_host, generator, outputs)


def lds_estep_backward(o, g, messages_out=False, message_seeds=None):
    """VJP w.r.t. (node_J, node_h): one launch of the reverse-mode kernel.

    messages_out: adjoint of the smoother / sampler ALONE (natural_smoother_general_grad + natural_sample_backward_grad,
    pyx:236-306,357-409): returns (g_M (B,T,n,n), g_h (B,T,n)), the raw cotangents of M_t = -2 J_filtered + J11_info and of
    the filtered h.  message_seeds = (g_M_pred, g_h_pred, g_M_filt, g_h_filt) (information form, any None): adjoint of the
    filter ALONE (natural_filter_grad, pyx:92-145) from cotangents of its messages (+ g["lognorm"]).

    `o` is the dict lds_estep_forward(..., save_for_bwd=True) returned (its workspace is consumed),
    `g` maps output names (samples, E_init_xx, E_init_x, E_pair_xx, E_pair_xxn, E_pair_xnxn, Exx_t, Exxn_t,
    E_node_xx, E_node_x, lognorm) to cotangent tensors (missing / None = zero).
    Mirrors natural_sample_backward_grad + natural_smoother_general_grad -> natural_filter_grad
    (cython_lds_inference.pyx:357-409, 236-306, 92-145; wiring lds_inference.py:26-39).
    """
    if "workspace" not in o or o["workspace"] is None:
        raise RuntimeError("lds_estep_backward: forward was not run with save_for_bwd=True (or already consumed)")
    B, T, n, S = o["B"], o["T"], o["n"], o["S"]
    dev = o["lognorm"].device
    gp = lambda k: None if g.get(k) is None else _lib.f64(g[k], dev)
    keep = {k: gp(k) for k in ("samples", "E_init_xx", "E_init_x", "E_pair_xx", "E_pair_xxn", "E_pair_xnxn",
                               "Exx_t", "Exxn_t", "E_node_xx", "E_node_x", "lognorm")}
    dense = bool(o.get("dense"))
    gJd = None
    if dense:     # cotangent of the full E[x_t x_t'] joins the per-time port; the result is a full matrix
        a, b2 = keep["E_node_xx"], keep["Exx_t"]
        keep["Exx_t"] = a if b2 is None else (b2 if a is None else a + b2)
        keep["E_node_xx"] = None
        gJd = torch.empty((B, T, n, n), dtype=torch.float64, device=dev)
    keep = {k: (None if v is None else v.contiguous()) for k, v in keep.items()}
    gJ = None if dense else torch.empty((B, T, n), dtype=torch.float64, device=dev)
    gh = torch.empty((B, T, n), dtype=torch.float64, device=dev)
    P = _lib.ptr
    flags, gM, gmh = 0, None, None
    seeds = (None,) * 4
    if messages_out:
        flags = _lib.SVAE_LDS_BWD_MESSAGES_OUT
        gM, gmh = torch.empty((B, T, n, n), dtype=torch.float64, device=dev), torch.empty((B, T, n), dtype=torch.float64, device=dev)
    elif message_seeds is not None:
        flags = _lib.SVAE_LDS_BWD_MESSAGES_IN
        seeds = tuple(None if x is None else _lib.f64(x, dev).contiguous() for x in message_seeds)
    code = _lib.lib().svae_lds_estep_bwd_ports(
        B, T, n, S, flags,
        P(o["E_node_x"]), P(o["samples"]), P(o["eps"]),
        P(keep["samples"]), P(keep["E_init_xx"]), P(keep["E_init_x"]),
        P(keep["E_pair_xx"]), P(keep["E_pair_xxn"]), P(keep["E_pair_xnxn"]),
        P(keep["Exx_t"]), P(keep["Exxn_t"]),
        P(keep["E_node_xx"]), P(keep["E_node_x"]), P(keep["lognorm"]),
        P(gJ), P(gh), P(gJd), P(gM), P(gmh), P(seeds[0]), P(seeds[1]), P(seeds[2]), P(seeds[3]),
        ctypes.c_void_p(o["workspace"].data_ptr()), o["workspace_bytes"], _lib.stream_ptr(dev))
    _lib.check(code, "svae_lds_estep_bwd_ports")
    o["workspace"] = None   # single use, like the reference's intermediates
    if messages_out:
        return gM, gmh
    if dense:               # d/dJ_node of a symmetric port: -2 sym(Pbar_t)
        gJ = -(gJd + gJd.transpose(-1, -2))
    return gJ, gh


_OUT_HOMOG = ("samples", "lognorm", "E_node_xx", "E_node_x", "E_init_xx", "E_init_x", "E_pair_xx", "E_pair_xxn", "E_pair_xnxn")
_OUT_PER_T = ("samples", "lognorm", "E_node_xx", "E_node_x", "E_init_xx", "E_init_x", "Exx_t", "Exxn_t")


class _LDSEStepFn(torch.autograd.Function):
    """Differentiable fused E-step: gradients flow to the node potentials only, exactly as the reference
    defines them (defgrad on argument 2 of the filter, lds_inference.py:27-29)."""

    @staticmethod
    def forward(ctx, nJ, nh, nlogZ, eps, init_params, pair_params, S, per_t):
        node = (nJ, nh) if nlogZ is None else (nJ, nh, nlogZ)
        o = lds_estep_forward(init_params, pair_params, node, S, eps, per_t_pair_stats=per_t, save_for_bwd=True)
        names = _OUT_PER_T if per_t else _OUT_HOMOG
        if o["dense"] and per_t:      # E_node_xx IS Exx_t: hand autograd the tensor once
            names = tuple(k for k in names if k != "E_node_xx")
        # Outputs must not be reachable from ctx except through save_for_backward: ctx -> output -> grad_fn -> ctx is
        # a reference cycle the collector cannot break (every forward's outputs + eps + workspace would stay alive).
        ctx.names, ctx.has_logZ = names, nlogZ is not None
        ctx.state = {k: o[k] for k in ("B", "T", "n", "S", "dense", "eps", "workspace", "workspace_bytes")}
        ctx.has_samples = o["samples"] is not None
        ctx.set_materialize_grads(False)
        outs = tuple(o[k] for k in ctx.names)
        if not ctx.has_samples:
            outs = (torch.empty(0, device=nh.device, dtype=torch.float64),) + outs[1:]
        ctx.save_for_backward(o["E_node_x"], outs[0], o["lognorm"])
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gs):
        g = dict(zip(ctx.names, gs))
        E_node_x, samples, lognorm = ctx.saved_tensors
        o = dict(ctx.state, E_node_x=E_node_x, samples=samples if ctx.has_samples else None, lognorm=lognorm)
        if not ctx.has_samples:
            g["samples"] = None
        B, T = o["B"], o["T"]
        try:
            gJ, gh = lds_estep_backward(o, g)
        finally:
            ctx.state = None      # single use, like the reference's intermediates; frees eps + workspace now
        g_logZ = None
        if ctx.has_logZ and g.get("lognorm") is not None:
            g_logZ = g["lognorm"][:, None].expand(B, T).contiguous()
        return gJ, gh, g_logZ, None, None, None, None, None


def _latent_dim(node_params):
    h = node_params[1]
    return int(h.shape[-1]) if hasattr(h, "shape") else int(torch.as_tensor(h).shape[-1])


def _estep(init_params, pair_params, node_params, S, eps, per_t, generator=None):
    """Dispatch: plain launch when nothing requires grad, autograd Function otherwise; latent dimensions above the
    compiled kernels' range run the same recursions as batched device linear algebra (lds_large.py)."""
    if _latent_dim(node_params) > _lib.SVAE_LDS_MAX_N:
        from . import lds_large
        return lds_large.estep(init_params, pair_params, node_params, S, eps, per_t, generator)
    needs_grad = torch.is_grad_enabled() and any(torch.is_tensor(x) and x.requires_grad for x in node_params)
    if not needs_grad:
        return lds_estep_forward(init_params, pair_params, node_params, S, eps, per_t_pair_stats=per_t,
                                 generator=generator)
    h_any = node_params[1]
    device = h_any.device
    nJ, nh, nlogZ, batched = _canonical_node_params(node_params, device)
    S = int(S) if S else 0
    B, T, n = nh.shape
    if eps is None and S > 0:
        eps = torch.randn((B, T, S, n), dtype=torch.float64, device=device, generator=generator)
    elif eps is not None:
        eps = _lib.f64(eps, device)
        eps = eps[None] if (not batched and eps.dim() == 3) else eps
    outs = _LDSEStepFn.apply(nJ, nh, nlogZ, eps, init_params, pair_params, S, per_t)
    names = _OUT_PER_T if per_t else _OUT_HOMOG
    dense = nJ.dim() == 4
    if dense and per_t:
        names = tuple(k for k in names if k != "E_node_xx")
    o = dict(zip(names, outs))
    if dense and per_t:
        o["E_node_xx"] = o["Exx_t"]
    o.update(B=B, T=T, n=n, S=S, batched=batched, eps=eps, dense=dense)
    if S == 0:
        o["samples"] = None
    return o


def _pack_stats(o, homog):
    B, T, n = o["B"], o["T"], o["n"]
    dev = o["lognorm"].device
    ones = lambda *s: torch.ones(s, dtype=torch.float64, device=dev)
    E_init = (o["E_init_xx"], o["E_init_x"], ones(B), ones(B))
    if homog:
        E_pair = (o["E_pair_xx"], o["E_pair_xxn"], o["E_pair_xnxn"], (T - 1.0) * ones(B))
    else:
        E_pair = (o["Exx_t"][:, :-1], o["Exxn_t"], o["Exx_t"][:, 1:], ones(B, T - 1))
    E_node = (o["E_node_xx"], o["E_node_x"], ones(B, T))
    stats = (E_init, E_pair, E_node)
    samples, lognorm = o["samples"], o["lognorm"]
    if not o["batched"]:
        sq = lambda x: x[0]
        stats = tuple(tuple(sq(x) for x in part) for part in stats)
        samples = None if samples is None else samples[0]
        lognorm = lognorm[0]
    return samples, stats, lognorm


def _is_homog(pair_params):
    x = pair_params[0]
    return (x.dim() if torch.is_tensor(x) else torch.as_tensor(x).dim()) == 2


# ------------------------------------------------------------------------------------------------
# reference-named drivers
# ------------------------------------------------------------------------------------------------

def natural_lds_inference_general(natparam, node_params, num_samples=1, eps=None, generator=None):
    """(samples, (E_init_stats, E_pair_stats, E_node_stats), lognorm)  -- lds_inference.py:185-202.

    samples (B,T,S,n) [or (T,S,n)]; E_init = (E[x0 x0'], E[x0], 1, 1); E_pair = (sum E[x_t x_t'],
    sum E[x_t x_{t+1}'], sum E[x_{t+1} x_{t+1}'], T-1) for homogeneous dynamics, per-time arrays
    otherwise (cython_lds_inference.pyx:198-210); E_node = (diag E[x_t x_t'], E[x_t], 1).
    `eps` (B,T,S,n) injects the sampler's standard-normal noise, eps[:,t] consumed at time t; the
    reference draws it internally in reverse time order (cython_lds_inference.pyx:333).
    """
    init_params, pair_params = natparam
    homog = _is_homog(pair_params)
    o = _estep(init_params, pair_params, node_params, num_samples, eps, not homog, generator)
    return _pack_stats(o, homog)


cython_natural_lds_inference_general = natural_lds_inference_general


def natural_lds_estep_general(natparam, node_params):
    """(lognorm, expected_stats) -- lds_inference.py:223-237."""
    init_params, pair_params = natparam
    homog = _is_homog(pair_params)
    o = _estep(init_params, pair_params, node_params, 0, None, not homog)
    _, stats, lognorm = _pack_stats(o, homog)
    return lognorm, stats


cython_natural_lds_estep_general = natural_lds_estep_general


def natural_lds_sample(natparam, node_params, num_samples=1, eps=None, generator=None):
    """samples only -- lds_inference.py:253-264."""
    init_params, pair_params = natparam
    if _latent_dim(node_params) > _lib.SVAE_LDS_MAX_N:
        from . import lds_large
        o = lds_large.estep(init_params, pair_params, node_params, num_samples, eps, False, generator, want_stats=False)
    else:
        o = lds_estep_forward(init_params, pair_params, node_params, num_samples, eps,
                              want_stats=False, generator=generator)
    return o["samples"] if o["batched"] else o["samples"][0]


cython_natural_lds_sample = natural_lds_sample


# ------------------------------------------------------------------------------------------------
# the reference's three primitives (lds_inference.py:18-39, cython_lds_inference.pyx:28,149,310), each one
# launch of the same fused kernel with the message ports of svae_lds_estep_fwd_ex open -- and each DIFFERENTIABLE
# on its own like the reference's primitive_with_aux wiring (lds_inference.py:26-39): the filter w.r.t. its node
# potentials given cotangents of its messages (natural_filter_grad), the smoother and the sampler w.r.t. the
# forward messages (natural_smoother_general_grad, natural_sample_backward_grad), through the sweep-2-only /
# sweep-1-only modes of the reverse-mode kernel (svae_lds_estep_bwd_ports).
# ------------------------------------------------------------------------------------------------

def _sym_info(gJ_nat):
    """cotangent of a natural-form precision message J -> cotangent of J_info = -2 J (symmetric part)."""
    return None if gJ_nat is None else -0.25 * (gJ_nat + gJ_nat.transpose(-1, -2))


class _FilterFn(torch.autograd.Function):
    """natural_filter_forward_general with natural_filter_grad as its VJP (defgrad on argument 2, lds_inference.py:27-29)."""

    @staticmethod
    def forward(ctx, nJ, nh, nlogZ, init_params, pair_params):
        node = (nJ, nh) if nlogZ is None else (nJ, nh, nlogZ)
        o = lds_estep_forward(init_params, pair_params, node, 0, None, want_stats=False, messages_out=True, save_for_bwd=True)
        J_filt = o["J_pred"] + (torch.diag_embed(nJ) if nJ.dim() == 3 else nJ)
        h_filt = o["h_pred"] + nh
        ctx.state = {k: o[k] for k in ("B", "T", "n", "S", "dense", "eps", "workspace", "workspace_bytes")}
        ctx.has_logZ = nlogZ is not None
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(o["E_node_x"], o["lognorm"])
        return o["J_pred"], o["h_pred"], J_filt, h_filt, o["lognorm"]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gJp, ghp, gJf, ghf, gln):
        E_node_x, lognorm = ctx.saved_tensors
        o = dict(ctx.state, E_node_x=E_node_x, samples=None, lognorm=lognorm)
        B, T = o["B"], o["T"]
        try:
            gJ, gh = lds_estep_backward(o, dict(lognorm=gln), message_seeds=(_sym_info(gJp), ghp, _sym_info(gJf), ghf))
        finally:
            ctx.state = None
        g_logZ = gln[:, None].expand(B, T).contiguous() if (ctx.has_logZ and gln is not None) else None
        return gJ, gh, g_logZ, None, None


def natural_filter_forward_general(init_params, pair_params, node_params):
    """((prediction_messages, filtered_messages), lognorm) -- cython_lds_inference.pyx:28-90.

    Messages are natural-parameter pairs (J (B,T,n,n), h (B,T,n)) [no B axis for unbatched node_params];
    the filtered message is the prediction message conditioned on the node potential (pyx:64-66).
    Differentiable w.r.t. the node potentials (natural_filter_grad, pyx:92-145)."""
    if _latent_dim(node_params) > _lib.SVAE_LDS_MAX_N:      # batched device linear algebra + autograd (lds_large.py)
        from . import lds_large
        msgs, lognorm, batched = lds_large.filter_forward(init_params, pair_params, node_params)
        if not batched:
            return tuple(tuple(x[0] for x in m) for m in msgs), lognorm[0]
        return msgs, lognorm
    h_any = node_params[1]
    device = h_any.device if torch.is_tensor(h_any) and h_any.is_cuda else torch.device("cuda", torch.cuda.current_device())
    nJ, nh, nlogZ, batched = _canonical_node_params(node_params, device)
    if torch.is_grad_enabled() and any(torch.is_tensor(x) and x.requires_grad for x in node_params):
        J_pred, h_pred, J_filt, h_filt, lognorm = _FilterFn.apply(nJ, nh, nlogZ, init_params, pair_params)
    else:
        o = lds_estep_forward(init_params, pair_params, (nJ, nh, nlogZ), 0, None, want_stats=False, messages_out=True)
        J_pred, h_pred, lognorm = o["J_pred"], o["h_pred"], o["lognorm"]
        J_filt = J_pred + (torch.diag_embed(nJ) if nJ.dim() == 3 else nJ)
        h_filt = h_pred + nh
    msgs = ((J_pred, h_pred), (J_filt, h_filt))
    if not batched:
        msgs = tuple(tuple(x[0] for x in m) for m in msgs)
        return msgs, lognorm[0]
    return msgs, lognorm


class _MessagesFn(torch.autograd.Function):
    """Smoother and/or sampler from given filtered messages; VJP w.r.t. the messages (defgrad on argument 0,
    lds_inference.py:31-39)."""

    @staticmethod
    def forward(ctx, J_filt, h_filt, eps, pair_params, S, per_t, want_stats):
        B, T, n = h_filt.shape
        dev = h_filt.device
        zero = torch.zeros((B, T, n), dtype=torch.float64, device=dev)
        init = (-0.5 * torch.eye(n, dtype=torch.float64, device=dev), zero[0, 0], 0.)   # overridden by the messages
        o = lds_estep_forward(init, pair_params, (zero, zero, None), S, eps, per_t_pair_stats=per_t, want_stats=want_stats,
                              from_messages=(J_filt, h_filt), save_for_bwd=True)
        names = ["samples"] if S else []
        if want_stats:
            names += [k for k in (_OUT_PER_T if per_t else _OUT_HOMOG) if k not in ("samples", "lognorm")]
        ctx.names = names
        ctx.state = {k: o[k] for k in ("B", "T", "n", "S", "dense", "eps", "workspace", "workspace_bytes")}
        ctx.set_materialize_grads(False)
        smp = o["samples"] if S else torch.empty(0, dtype=torch.float64, device=dev)
        ctx.has_samples = bool(S)
        ctx.save_for_backward(o["E_node_x"], smp, o["lognorm"])
        return tuple(o[k] for k in names)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gs):
        g = dict(zip(ctx.names, gs))
        E_node_x, smp, lognorm = ctx.saved_tensors
        o = dict(ctx.state, E_node_x=E_node_x, samples=smp if ctx.has_samples else None, lognorm=lognorm)
        try:
            gM, gmh = lds_estep_backward(o, g, messages_out=True)
        finally:
            ctx.state = None
        # M_t = -2 J_filtered + J11_info: d/dJ_filtered (natural form, symmetric port) = -2 sym(d/dM)
        return -(gM + gM.transpose(-1, -2)), gmh, None, None, None, None, None


def _messages_run(forward_messages, pair_params, num_samples, eps, generator, want_stats):
    _, (J_filt, h_filt) = forward_messages
    device = h_filt.device if torch.is_tensor(h_filt) and h_filt.is_cuda else torch.device("cuda", torch.cuda.current_device())
    needs_grad = torch.is_grad_enabled() and any(torch.is_tensor(x) and x.requires_grad for x in (J_filt, h_filt))
    J_filt, h_filt = _lib.f64(J_filt, device), _lib.f64(h_filt, device)
    batched = h_filt.dim() == 3
    if not batched:
        J_filt, h_filt = J_filt[None], h_filt[None]
        if eps is not None and torch.as_tensor(eps).dim() == 3:
            eps = torch.as_tensor(eps)[None]
    B, T, n = h_filt.shape
    homog = _is_homog(pair_params)
    S = int(num_samples) if num_samples else 0
    if n > _lib.SVAE_LDS_MAX_N:                              # batched device linear algebra + autograd (lds_large.py)
        from . import lds_large
        o = lds_large.from_messages(J_filt, h_filt, pair_params, S, eps, not homog, want_stats, generator)
        o["batched"] = batched
        return o, homog
    if needs_grad:
        if eps is None and S > 0:
            eps = torch.randn((B, T, S, n), dtype=torch.float64, device=device, generator=generator)
        elif eps is not None:
            eps = _lib.f64(eps, device).contiguous()
        outs = _MessagesFn.apply(J_filt.contiguous(), h_filt.contiguous(), eps, pair_params, S, not homog, want_stats)
        names = (["samples"] if S else []) + ([k for k in (_OUT_HOMOG if homog else _OUT_PER_T) if k not in ("samples", "lognorm")]
                                              if want_stats else [])
        o = dict(zip(names, outs))
        o.update(B=B, T=T, n=n, S=S, dense=False, eps=eps, lognorm=torch.zeros(B, dtype=torch.float64, device=device))
        if S == 0:
            o["samples"] = None
    else:
        zero = torch.zeros((B, T, n), dtype=torch.float64, device=device)
        init = (-0.5 * torch.eye(n, dtype=torch.float64, device=device), zero[0, 0], 0.)   # overridden by the messages
        o = lds_estep_forward(init, pair_params, (zero, zero, None), S, eps, per_t_pair_stats=not homog,
                              want_stats=want_stats, generator=generator, from_messages=(J_filt, h_filt))
    o["batched"] = batched
    return o, homog


def natural_smoother_general(forward_messages, pair_params):
    """(E_init_stats, E_pair_stats, E_node_stats) from given forward messages -- pyx:149-195.
    Differentiable w.r.t. the filtered messages (natural_smoother_general_grad, pyx:236-306); the covariance-form
    recursion does not read the prediction messages, whose gradient is zero (see include/svae_b200.h)."""
    o, homog = _messages_run(forward_messages, pair_params, 0, None, None, True)
    return _pack_stats(o, homog)[1]


def natural_sample_backward(forward_messages, pair_params, num_samples=1, eps=None, generator=None):
    """samples (B,T,S,n) drawn backwards from given filtered messages -- pyx:310-355.
    Differentiable w.r.t. the filtered messages (natural_sample_backward_grad, pyx:357-409)."""
    o, _ = _messages_run(forward_messages, pair_params, num_samples, eps, generator, False)
    return o["samples"] if o["batched"] else o["samples"][0]


cython_natural_filter_forward = natural_filter_forward_general
cython_natural_smoother_general = natural_smoother_general
cython_natural_sample_backward = natural_sample_backward
