"""SLDS-SVAE composite local E-step -- torch mirror of the reference's svae/models/slds_svae.py:80-310.

Block mean field alternating the HMM E-step (csrc/hmm_estep.cu) and the LDS E-step with time-inhomogeneous,
per-sequence pair parameters (csrc/lds_estep.cuh, INHOMOG), then one differentiable LDS pass + hmm_logZ
(slds_svae.py:289-310).  Batched over B sequences: node potentials (B,T,n); the discrete/continuous factors of
different sequences are independent, global statistics are summed over B.  With B = 1 (or unbatched (T,n)
input) this is the reference's algorithm step for step; with B > 1 every sequence stops on its OWN bound (its factors
are frozen from then on), so a batch reproduces B independent reference runs.

Global natural parameters: (hmm_global, lds_global) with hmm_global = (dir_natparam (K,), mdir_natparam (K,K))
and lds_global = [(niw_natparam packed (n+2,n+2), mniw_natparam (A,B,C,d))] * K  (slds_svae.py:26-75).
"""
from __future__ import annotations
import sys
import torch
from ..util import sub, contract, unbox
from ..lds.lds_inference import natural_lds_inference_general, natural_lds_estep_general, natural_lds_sample
from ..hmm.hmm_inference import hmm_estep, hmm_logZ
from ..distributions import niw, mniw, dirichlet, gaussian


# ---- global natural parameters (slds_svae.py:26-75) ---------------------------------------------------

def make_hmm_global_natparam(num_states, alpha=1., sticky_bias=0., random=False, generator=None, device=None):
    mk = lambda: (alpha * torch.ones(num_states, dtype=torch.float64) if not random
                  else alpha + torch.rand(num_states, dtype=torch.float64, generator=generator))
    dir_natparam = mk()
    mdir_natparam = sticky_bias * torch.eye(num_states, dtype=torch.float64) + torch.stack([mk() for _ in range(num_states)])
    return dir_natparam.to(device), mdir_natparam.to(device)


def make_lds_global_natparams(num_states, state_dim, random=False, generator=None, device=None):
    n = state_dim
    eye = torch.eye(n, dtype=torch.float64)
    r = lambda *s: torch.rand(*s, dtype=torch.float64, generator=generator)
    rn = lambda *s: torch.randn(*s, dtype=torch.float64, generator=generator)
    out = []
    for _ in range(num_states):
        if not random:
            nu, S, mu, kappa = n + 10., (n + 10.) * eye, torch.zeros(n, dtype=torch.float64), 10.
            nu2, S2, M, K = n + 10., (n + 10.) * eye, torch.zeros(n, n, dtype=torch.float64), 10. * eye
        else:
            nu, S, mu, kappa = n + 4. + float(r(())), (n + float(r(()))) * eye, rn(n), float(r(()))
            nu2, S2, M, K = n + 4. + float(r(())), (n + float(r(()))) * eye, 1e-2 * rn(n, n), (n + float(r(()))) * eye
        out.append((niw.standard_to_natural(S, mu, kappa, nu).to(device),
                    tuple(x.to(device) for x in mniw.standard_to_natural(nu2, S2, M, K))))
    return out


def make_slds_global_natparam(num_states, state_dim, alpha=5., sticky_bias=0., random=False, generator=None, device=None):
    return (make_hmm_global_natparam(num_states, alpha, sticky_bias, random, generator, device),
            make_lds_global_natparams(num_states, state_dim, random, generator, device))


# ---- per-mode local parameters ----------------------------------------------------------------------

def stack_lds_global(lds_global_natparams):
    """[(niw (n+2,n+2), (A, B, C, d))] * K  ->  niw (K,n+2,n+2), (A (K,n,n), B, C, d (K,)): the per-mode exp-family maps
    then run batched over the modes (a handful of launches instead of ~40 per mode)."""
    if isinstance(lds_global_natparams, StackedLdsGlobal):
        return lds_global_natparams
    niw_nat = torch.stack([a for a, _ in lds_global_natparams])
    mniw_nat = tuple(torch.stack([torch.as_tensor(b[j], dtype=niw_nat.dtype, device=niw_nat.device) for _, b in lds_global_natparams])
                     for j in range(4))
    return StackedLdsGlobal((niw_nat, mniw_nat))


class StackedLdsGlobal(tuple):
    """(niw_stack, mniw_stack) marker type."""


def lds_global_expectedstats(lds_global_natparams):
    """(niw.expectedstats (K,n+2,n+2) packed, mniw.expectedstats 4-tuple of (K,...)) -- models/lds.py:22-24 for all modes."""
    niw_nat, mniw_nat = stack_lds_global(lds_global_natparams)
    return niw.expectedstats(niw_nat), mniw.expectedstats(mniw_nat)


def get_all_lds_local_natparams(lds_global_natparams, expected=None):
    """slds_svae.py:86-89: dense stacks over the K modes: init (J (K,n,n), h (K,n), a (K,), b (K,)) and
    pair (J11, J12, J22 (K,n,n), logZ (K,))."""
    e_niw, e_mniw = expected if expected is not None else lds_global_expectedstats(lds_global_natparams)
    return tuple(gaussian.unpack_dense(e_niw)), tuple(e_mniw)


def get_var_lds_local_natparam(lds_global_natparams, hmm_expected_stats, dense=None):
    """slds_svae.py:92-103: mix the per-mode parameters with E[z_t] -> per-sequence init, time-inhomogeneous pair."""
    _, _, Ez = hmm_expected_stats                                     # (B,T,K)
    dense_init, dense_pair = dense if dense is not None else get_all_lds_local_natparams(lds_global_natparams)
    B, T, K = Ez.shape

    def mix(w, p):   # (R,K) @ (K,M): a plain tall-skinny GEMM (cuBLAS), reshaped back
        return (w.reshape(-1, K) @ p.reshape(K, -1)).reshape(w.shape[:-1] + p.shape[1:])
    init_param = tuple(mix(Ez[:, 0], p) for p in dense_init)          # (B,n,n), (B,n), (B,), (B,)
    w = Ez[:, 1:].contiguous()
    pair_params = tuple(mix(w, p) for p in dense_pair)                # (B,T-1,n,n) x3, (B,T-1)
    return init_param, pair_params


def use_parallel_in_time(B, T, flag="auto"):
    """Few long sequences leave the sequential kernels on B tiles of the GPU: chunk-parallel drivers then
    (lds/lds_pit.py, hmm/hmm_pit.py; same numbers).  Measured crossover on a B200: B <= 16 and T >= 2000."""
    if flag is None or flag == "auto":
        return B <= 16 and T >= 2000
    return bool(flag)


def lds_meanfield(lds_global_natparams, node_potentials, hmm_expected_stats, dense=None, pit=False):
    """slds_svae.py:80-83."""
    local_natparam = get_var_lds_local_natparam(lds_global_natparams, hmm_expected_stats, dense)
    if pit:
        from ..lds.lds_pit import natural_lds_inference_parallel
        _, expected_stats, vlb = natural_lds_inference_parallel(local_natparam, node_potentials, 0)
    else:
        vlb, expected_stats = natural_lds_estep_general(local_natparam, node_potentials)
    return local_natparam, expected_stats, vlb


def hmm_prior_expectedstats(natparam):
    """slds_svae.py:122-128."""
    dir_natparam, mdir_natparam = natparam
    return dirichlet.expectedstats(dir_natparam), dirichlet.expectedstats(mdir_natparam)


get_hmm_local_natparam = hmm_prior_expectedstats


def get_arhmm_local_nodeparams(lds_global_natparam, lds_expected_stats, dense=None):
    """slds_svae.py:131-147: HMM node potential (t,k) = <expected LDS statistics at t, mode-k parameters>."""
    init_stats, pair_stats = lds_expected_stats[:2]
    (J, h, a, b), (J11, J12, J22, logZ) = dense if dense is not None else get_all_lds_local_natparams(lds_global_natparam)
    ExxT0, Ex0 = init_stats[0], init_stats[1]
    K = J.shape[0]
    flatK = lambda p: p.reshape(K, -1).T                              # (M,K)
    init_node = ExxT0.reshape(ExxT0.shape[0], -1) @ flatK(J) + Ex0 @ h.T + (a + b)[None]
    Exx, Exxn, Exnxn = pair_stats[0], pair_stats[1], pair_stats[2]    # (B,T-1,n,n)
    B, Tm = Exx.shape[0], Exx.shape[1]
    # three tall-skinny GEMMs (R, n^2) @ (n^2, K) instead of einsum's batched path
    rest = (Exx.reshape(B * Tm, -1) @ flatK(J11) + Exxn.reshape(B * Tm, -1) @ flatK(J12)
            + Exnxn.reshape(B * Tm, -1) @ flatK(J22)).reshape(B, Tm, K) + logZ[None, None]
    return torch.cat((init_node[:, None], rest), 1)                   # (B,T,K)


def hmm_meanfield(hmm_global_natparam, lds_global_natparam, lds_expected_stats, dense=None, hmm_local=None, pit=False):
    """slds_svae.py:108-115."""
    init_params, pair_params = hmm_local if hmm_local is not None else get_hmm_local_natparam(hmm_global_natparam)
    node_params = get_arhmm_local_nodeparams(lds_global_natparam, lds_expected_stats, dense)
    local_natparam = (init_params, pair_params, node_params)
    if pit:
        from ..hmm.hmm_pit import hmm_estep_parallel
        vlb, stats = hmm_estep_parallel(local_natparam)
    else:
        vlb, stats = hmm_estep(local_natparam)
    return local_natparam, stats, vlb


def get_hmm_vlb(lds_global_natparam, hmm_local_natparam, lds_expected_stats, dense=None):
    """slds_svae.py:150-154."""
    init_params, pair_params, _ = hmm_local_natparam
    node_params = get_arhmm_local_nodeparams(lds_global_natparam, lds_expected_stats, dense)
    return hmm_logZ((init_params, pair_params, node_params))


def initialize_sample_path(node_potentials, eps=None, generator=None, pit=False):
    """slds_svae.py:203-223: ONE sample path (B,T,n) of a 0.9-random-walk LDS given the node potentials."""
    h = node_potentials[1]
    N, dev = h.shape[-1], h.device
    eye = torch.eye(N, dtype=torch.float64, device=dev)
    A = 0.9 * eye
    init_params = (-0.5 * eye, torch.zeros(N, dtype=torch.float64, device=dev), 0.)
    pair_params = (-0.5 * A.T @ A, A.T, -0.5 * eye, 0.)
    if pit:
        from ..lds.lds_pit import natural_lds_inference_parallel
        return natural_lds_inference_parallel((init_params, pair_params), node_potentials, 1, eps=eps, generator=generator)[0][:, :, 0]
    return natural_lds_sample((init_params, pair_params), node_potentials, num_samples=1, eps=eps, generator=generator)[:, :, 0]


def initialize_local_meanfield(node_potentials, eps=None, generator=None, pit=False, path=None):
    """slds_svae.py:203-226: the statistics of that sample path (compute_stats, :206-215)."""
    x = path if path is not None else initialize_sample_path(node_potentials, eps, generator, pit)
    dev = x.device
    out = lambda u, v: u[..., :, None] * v[..., None, :]
    ones = torch.ones(x.shape[0], dtype=torch.float64, device=dev)
    init_stats = (out(x[:, 0], x[:, 0]), x[:, 0], ones, ones)
    pair_stats = (out(x[:, :-1], x[:, :-1]), out(x[:, :-1], x[:, 1:]), out(x[:, 1:], x[:, 1:]),
                  torch.ones(x.shape[0], x.shape[1] - 1, dtype=torch.float64, device=dev))
    return init_stats, pair_stats


def _tree_where(mask, new, old):
    """Per-sequence select over nested tuples of (B, ...) tensors."""
    if isinstance(new, (tuple, list)):
        return tuple(_tree_where(mask, a, b) for a, b in zip(new, old))
    if not torch.is_tensor(new) or new.dim() == 0 or new.shape[0] != mask.shape[0]:
        return new
    return torch.where(mask.reshape((-1,) + (1,) * (new.dim() - 1)), new, old)


def optimize_local_meanfield(global_natparam, node_potentials, tol=1e-2, max_iter=100, init_eps=None, generator=None,
                             dense=None, parallel_in_time="auto", hmm_local=None):
    """slds_svae.py:159-175, batched: EVERY sequence follows the reference's rule on its own bound -- it stops (its
    factors are frozen) after the first sweep that moved its bound by less than `tol`, so B sequences in one call
    reproduce B reference calls whatever their sweep counts.  The per-mode expected parameters (loop-invariant) are
    evaluated once per call, batched over the K modes.

    Default path (`_meanfield_device`): a sweep is four launches on persistent buffers -- svae_slds_node_potentials,
    the HMM kernel, svae_slds_mix_params, the fused LDS kernel (inhomogeneous, per-sequence) -- all taking the (B,)
    `active` mask, so converged sequences cost nothing and keep their factors; one small host read per sweep decides
    whether any sequence is still active.  parallel_in_time ("auto" | True | False, use_parallel_in_time) selects the
    chunk-parallel drivers for few long sequences (torch glue, `_meanfield_torch`)."""
    hmm_global, lds_global = global_natparam
    dense = dense if dense is not None else get_all_lds_local_natparams(lds_global)
    hmm_local = hmm_local if hmm_local is not None else get_hmm_local_natparam(hmm_global)
    pit = use_parallel_in_time(node_potentials[1].shape[0], node_potentials[1].shape[1], parallel_in_time) \
        and not (torch.is_grad_enabled() and any(p.requires_grad for p in node_potentials[:2]))
    path = initialize_sample_path(node_potentials, init_eps, generator, pit)
    B, T, n = node_potentials[1].shape
    K = hmm_local[0].shape[0]
    if not pit and K <= 32 and 8 * (4 * K * n * n + K * n + 2 * K) <= 200 * 1024:
        return _meanfield_device(hmm_local, dense, node_potentials, path, tol, max_iter)
    lds_stats = initialize_local_meanfield(node_potentials, path=path)
    return _meanfield_torch(hmm_global, lds_global, dense, hmm_local, node_potentials, lds_stats, tol, max_iter, pit)


def _meanfield_device(hmm_local, dense, node_potentials, path, tol, max_iter):
    import ctypes
    from .. import _lib
    from ..lds.lds_inference import lds_estep_forward
    from ..hmm.hmm_inference import hmm_estep_forward
    nJ, nh = node_potentials[0].detach().contiguous(), node_potentials[1].detach().contiguous()
    dev = nh.device
    B, T, n = nh.shape
    log_pi, log_A = (_lib.f64(x, dev).contiguous() for x in hmm_local)
    K = log_pi.shape[0]
    (iJ, ih, ia, ib), (J11, J12, J22, lZ) = dense
    iJ, ih, J11, J12, J22, lZ = (_lib.f64(x, dev).contiguous() for x in (iJ, ih, J11, J12, J22, lZ))
    ic = (_lib.f64(ia, dev) + _lib.f64(ib, dev)).contiguous()
    new = lambda *s_: torch.empty(s_, dtype=torch.float64, device=dev)
    Tm = max(T - 1, 1)
    # per-time second moments the node-potential kernel reads: first those of the initial sample path (compute_stats,
    # slds_svae.py:206-215), then the LDS kernel rewrites them in place every sweep
    Exx_t, Exxn_t = new(B, T, n, n), new(B, Tm, n, n)
    torch.mul(path[..., :, None], path[..., None, :], out=Exx_t)
    if T > 1:
        torch.mul(path[:, :-1, :, None], path[:, 1:, None, :], out=Exxn_t)
    buf = dict(Exx_t=Exx_t, Exxn_t=Exxn_t if T > 1 else None, E_init_xx=new(B, n, n), E_init_x=path[:, 0].contiguous().clone(),
               E_node_xx=new(B, T, n), E_node_x=new(B, T, n), lognorm=new(B),
               E_pair_xx=new(B, n, n), E_pair_xxn=new(B, n, n), E_pair_xnxn=new(B, n, n),
               status=torch.zeros(1, dtype=torch.int32, device=dev),
               workspace=torch.empty(max(_lib.lib().svae_lds_estep_workspace_bytes(B, T, n, 0, 0), 16), dtype=torch.uint8, device=dev))
    node = new(B, T, K)
    hmm_out = (new(B), new(B, K), new(B, K, K), new(B, T, K))
    m_iJ, m_ih, m_ic = new(B, n, n), new(B, n), new(B)
    m11, m12, m22, mlZ = new(B, Tm, n, n), new(B, Tm, n, n), new(B, Tm, n, n), new(B, Tm)
    active = torch.ones(B, dtype=torch.int32, device=dev)
    L, P, st = _lib.lib(), _lib.ptr, _lib.stream_ptr(dev)
    ap = ctypes.c_void_p(active.data_ptr())
    prev = None
    for _ in range(max_iter):
        _lib.check(L.svae_slds_node_potentials(B, T, n, K, P(buf["Exx_t"]), P(Exxn_t), P(buf["E_init_x"]), P(iJ), P(ih), P(ic),
                                               P(J11), P(J12), P(J22), P(lZ), ap, P(node), st), "svae_slds_node_potentials")
        hmm_estep_forward(log_pi, log_A, node, active=active, out=hmm_out)
        _lib.check(L.svae_slds_mix_params(B, T, n, K, P(hmm_out[3]), P(iJ), P(ih), P(ic), P(J11), P(J12), P(J22), P(lZ), ap,
                                          P(m_iJ), P(m_ih), P(m_ic), P(m11), P(m12), P(m22), P(mlZ), st), "svae_slds_mix_params")
        pair = (m11, m12, m22, mlZ) if T > 1 else (m11[:, :0], m12[:, :0], m22[:, :0], mlZ[:, :0])
        o = lds_estep_forward((m_iJ, m_ih, m_ic), pair, (nJ, nh), 0, per_t_pair_stats=True, buffers=buf, active=active)
        vlb = hmm_out[0] + o["lognorm"]
        if prev is not None:
            active.mul_(((vlb - prev).abs() >= tol).to(torch.int32))       # frozen sequences: vlb == prev, stay 0
            if not bool(active.any()):                                    # the one host read per sweep
                break
        prev = vlb.clone()
    else:
        print('iteration limit reached', file=sys.stderr)
    ones = lambda *s_: torch.ones(s_, dtype=torch.float64, device=dev)
    hmm_vlb, Ez0, Etrans, Ez = hmm_out
    E_init = (o["E_init_xx"], o["E_init_x"], ones(B), ones(B))
    E_pair = (o["Exx_t"][:, :-1], o["Exxn_t"], o["Exx_t"][:, 1:], ones(B, T - 1)) if T > 1 else \
        tuple(new(B, 0, n, n) for _ in range(3)) + (ones(B, 0),)
    E_node = (o["E_node_xx"], o["E_node_x"], ones(B, T))
    hmm_natparam = (log_pi, log_A, node)
    lds_natparam = ((m_iJ, m_ih, m_ic), pair)
    return ((Ez0, Etrans, Ez), (E_init, E_pair, E_node)), (hmm_natparam, lds_natparam), (hmm_vlb, o["lognorm"])


def _meanfield_torch(hmm_global, lds_global, dense, hmm_local, node_potentials, lds_stats, tol, max_iter, pit):
    """The same loop in torch glue around the E-step launches (parallel-in-time drivers; shapes beyond the glue kernels)."""
    B = node_potentials[1].shape[0]
    active = torch.ones(B, dtype=torch.bool, device=node_potentials[1].device)
    cur, prev_vlb = None, None
    for _ in range(max_iter):
        hmm_natparam, hmm_stats, hmm_vlb = hmm_meanfield(hmm_global, lds_global, lds_stats, dense, hmm_local, pit)
        lds_natparam, lds_stats_new, lds_vlb = lds_meanfield(lds_global, node_potentials, hmm_stats, dense, pit)
        new = (hmm_natparam, hmm_stats, hmm_vlb, lds_natparam, lds_stats_new, lds_vlb)
        cur = new if cur is None else _tree_where(active, new, cur)
        vlb = cur[2] + cur[5]
        if prev_vlb is not None:
            active = active & ~((vlb - prev_vlb).abs() < tol)
            if not bool(active.any()):                      # the one host sync per sweep
                break
        prev_vlb = vlb
        lds_stats = cur[4]
    else:
        print('iteration limit reached', file=sys.stderr)
    hmm_natparam, hmm_stats, hmm_vlb, lds_natparam, lds_stats, lds_vlb = cur
    return (hmm_stats, lds_stats), (hmm_natparam, lds_natparam), (hmm_vlb, lds_vlb)


def get_global_stats(hmm_stats, lds_stats):
    """slds_svae.py:229-243, summed over the batch; same structure as the global natural parameters.
    One (K, B T) x (B T, n^2) product per statistic instead of a loop over the modes."""
    Ez0, Etrans, Ez = hmm_stats
    init_stats, pair_stats = lds_stats
    hmm_global = (Ez0.sum(0), Etrans.sum(0))
    K = Ez.shape[-1]
    w0 = Ez[:, 0]                                                     # (B,K)
    wt = Ez[:, 1:].reshape(-1, K)                                     # (B(T-1),K)
    gi = tuple((w0.T @ s_.reshape(s_.shape[0], -1)).reshape((K,) + tuple(s_.shape[1:])) for s_ in init_stats)
    gp = tuple((wt.T @ s_.reshape(wt.shape[0], -1)).reshape((K,) + tuple(s_.shape[2:])) for s_ in pair_stats)
    packed = gaussian.pack_dense(gi[0], gi[1], gi[2], gi[3])          # (K,n+2,n+2)
    lds_global = [(packed[k], tuple(x[k] for x in gp)) for k in range(K)]
    return hmm_global, lds_global


def slds_prior_expectedstats(global_natparam, lds_expected=None):
    """slds_svae.py:272-284."""
    hmm_natparam, lds_natparams = global_natparam
    e_niw, e_mniw = lds_expected if lds_expected is not None else lds_global_expectedstats(lds_natparams)
    K = e_niw.shape[0]
    return hmm_prior_expectedstats(hmm_natparam), [(e_niw[k], tuple(x[k] for x in e_mniw)) for k in range(K)]


def slds_prior_logZ(natparam):
    """slds_svae.py:254-269 (the per-mode terms batched over K)."""
    (dir_nat, mdir_nat), lds_prior_natparams = natparam
    niw_nat, mniw_nat = stack_lds_global(lds_prior_natparams)
    return dirichlet.logZ(dir_nat) + dirichlet.logZ(mdir_nat) + niw.logZ(niw_nat) + mniw.logZ(mniw_nat)


def _tuplify(x):
    return tuple(_tuplify(y) for y in x) if isinstance(x, (list, tuple)) else x


def slds_prior_vlb(global_natparam, prior_natparam, lds_expected=None):
    """slds_svae.py:248-251: <prior - global, E_global[stats]> - (logZ(prior) - logZ(global)), evaluated on the stacked
    per-mode tensors (the inner product of the nested structures is the sum of the component inner products)."""
    (g_dir, g_mdir), g_lds = global_natparam
    (p_dir, p_mdir), p_lds = prior_natparam
    g_niw, g_mniw = stack_lds_global(g_lds)
    p_niw, p_mniw = stack_lds_global(p_lds)
    e_niw, e_mniw = lds_expected if lds_expected is not None else lds_global_expectedstats(g_lds)
    e_dir, e_mdir = hmm_prior_expectedstats((g_dir, g_mdir))
    inner = ((p_dir - g_dir) * e_dir).sum() + ((p_mdir - g_mdir) * e_mdir).sum() + ((p_niw - g_niw) * e_niw).sum()
    for pa, ga, ea in zip(p_mniw, g_mniw, e_mniw):
        inner = inner + ((pa - ga) * ea).sum()
    return inner - (slds_prior_logZ(prior_natparam) - slds_prior_logZ(global_natparam))



# ---- the global side of a step, replayed as ONE CUDA graph ------------------------------------------------
# lds_global_expectedstats + get_all_lds_local_natparams + slds_prior_vlb are ~300 launches on (K, n+2, n+2)-sized
# tensors: host-bound (about 4 ms of an SVI step whose kernels take 6).  They are a pure function of the stacked global
# and prior natural parameters with static shapes, so the sequence is captured once per (shapes, device) and replayed:
# inputs are copied into the graph's static buffers, outputs are read from them (the caller gets clones of what it
# returns to its own caller).  Only taken when nothing on the global side needs a gradient -- the SVI step takes the
# natural gradient in closed form (svae.py:33-34) and never differentiates the global KL -- and SVAE_SLDS_GRAPH != 0.

def _global_side(hmm_global, g_niw, g_mniw, hmm_prior, p_niw, p_mniw):
    lds_global = StackedLdsGlobal((g_niw, tuple(g_mniw)))
    lds_expected = lds_global_expectedstats(lds_global)
    dense = get_all_lds_local_natparams(lds_global, lds_expected)
    global_vlb = slds_prior_vlb((tuple(hmm_global), lds_global), (tuple(hmm_prior), StackedLdsGlobal((p_niw, tuple(p_mniw)))),
                                lds_expected)
    hmm_local = get_hmm_local_natparam(tuple(hmm_global))
    return lds_expected, dense, global_vlb, hmm_local


_GLOBAL_GRAPHS = {}


def _flatten(x, out):
    if isinstance(x, (tuple, list)):
        for y in x:
            _flatten(y, out)
    else:
        out.append(x)
    return out


def _unflatten_like(x, it):
    if isinstance(x, (tuple, list)):
        return tuple(_unflatten_like(y, it) for y in x)
    return next(it)


def global_side(hmm_global, lds_global, hmm_prior, lds_prior):
    """(lds_expected, dense, global_vlb, hmm_local) for stacked global / prior natural parameters; graph replay on CUDA."""
    import os
    args = (tuple(hmm_global), lds_global[0], tuple(lds_global[1]), tuple(hmm_prior), lds_prior[0], tuple(lds_prior[1]))
    flat_in = _flatten(args, [])
    eligible = (os.environ.get("SVAE_SLDS_GRAPH", "1") != "0" and all(torch.is_tensor(t) and t.is_cuda for t in flat_in)
                and not (torch.is_grad_enabled() and any(t.requires_grad for t in flat_in))
                and not torch.cuda.is_current_stream_capturing())
    if not eligible:
        return _global_side(*args)
    key = (flat_in[0].device,) + tuple((tuple(t.shape), t.dtype) for t in flat_in)
    entry = _GLOBAL_GRAPHS.get(key)
    if entry is None:
        entry = _GLOBAL_GRAPHS[key] = _capture_global_side(args, flat_in)
    if entry is False:                                   # capture was refused once (an op that synchronises): stay eager
        return _global_side(*args)
    graph, static_in, static_out = entry
    with torch.no_grad():
        torch._foreach_copy_(static_in, [t.detach() for t in flat_in])
        graph.replay()
    lds_expected, dense, global_vlb, hmm_local = static_out
    # static buffers are overwritten by the next call (silently: a replay does not bump version counters), so whatever
    # can outlive this step -- returned values, tensors the autograd graph of the differentiable pass saves -- is cloned
    dense = tuple(tuple(x.clone() for x in grp) for grp in dense)
    return lds_expected, dense, global_vlb.clone(), tuple(x.clone() for x in hmm_local)


def _capture_global_side(args, flat_in):
    try:
        with torch.no_grad():
            static_in = [t.detach().clone() for t in flat_in]
            static_args = _unflatten_like(args, iter(static_in))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                        # warm-up off the capture (lazy handles, workspaces)
                    _global_side(*static_args)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = _global_side(*static_args)
        return graph, static_in, static_out
    except Exception as e:                                # noqa: BLE001 -- any capture failure means "run eagerly"
        print("svae_b200: SLDS global-side graph capture refused (%s: %s); running it eagerly" % (type(e).__name__, e), file=sys.stderr)
        torch.cuda.synchronize()
        return False


def run_inference(prior_natparam, global_natparam, nn_potentials, num_samples, eps=None, init_eps=None, generator=None):
    """slds_svae.py:289-310: returns (samples, expected_stats, global_vlb, local_vlb)."""
    hmm_global_natparam, lds_global_natparam = global_natparam
    batched = nn_potentials[1].dim() == 3
    pots = tuple(nn_potentials[:2]) if batched else tuple(p[None] for p in nn_potentials[:2])
    # the K modes stacked ONCE (global and prior), every per-mode map batched over them and reused by the prior term
    lds_global_natparam = stack_lds_global(lds_global_natparam)
    global_natparam = (hmm_global_natparam, lds_global_natparam)
    prior_natparam = (prior_natparam[0], stack_lds_global(prior_natparam[1]))
    lds_expected, dense, global_vlb, hmm_local = global_side(hmm_global_natparam, lds_global_natparam, *prior_natparam)
    with torch.no_grad():
        (hmm_stats, _), (hmm_local_natparam, lds_local_natparam), _ = \
            optimize_local_meanfield(global_natparam, unbox(pots), init_eps=init_eps, generator=generator, dense=dense,
                                     hmm_local=hmm_local)
    samples, lds_stats, lds_normalizer = natural_lds_inference_general(lds_local_natparam, pots, num_samples, eps=eps,
                                                                       generator=generator)
    hmm_vlb = get_hmm_vlb(lds_global_natparam, hmm_local_natparam, lds_stats, dense)
    global_lds_stats, local_lds_stats = lds_stats[:-1], lds_stats[-1]
    expected_stats = get_global_stats(hmm_stats, global_lds_stats)
    lds_vlb = lds_normalizer.sum() - contract(pots, local_lds_stats[:2])
    local_vlb = hmm_vlb.sum() + lds_vlb
    if not batched:
        samples = samples[0]
    return samples, expected_stats, global_vlb, local_vlb
