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

def get_all_lds_local_natparams(lds_global_natparams):
    """slds_svae.py:86-89: dense stacks over the K modes: init (J (K,n,n), h (K,n), a (K,), b (K,)) and
    pair (J11, J12, J22 (K,n,n), logZ (K,))."""
    inits, pairs = [], []
    for niw_nat, mniw_nat in lds_global_natparams:
        inits.append(gaussian.unpack_dense(niw.expectedstats(niw_nat)))
        pairs.append(mniw.expectedstats(mniw_nat))
    dense_init = tuple(torch.stack(x) for x in zip(*inits))
    dense_pair = tuple(torch.stack(x) for x in zip(*pairs))
    return dense_init, dense_pair


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


def initialize_local_meanfield(node_potentials, eps=None, generator=None, pit=False):
    """slds_svae.py:203-226: statistics of ONE sample path of a 0.9-random-walk LDS given the node potentials."""
    h = node_potentials[1]
    N, dev = h.shape[-1], h.device
    eye = torch.eye(N, dtype=torch.float64, device=dev)
    A = 0.9 * eye
    init_params = (-0.5 * eye, torch.zeros(N, dtype=torch.float64, device=dev), 0.)
    pair_params = (-0.5 * A.T @ A, A.T, -0.5 * eye, 0.)
    if pit:
        from ..lds.lds_pit import natural_lds_inference_parallel
        x = natural_lds_inference_parallel((init_params, pair_params), node_potentials, 1, eps=eps, generator=generator)[0][:, :, 0]
    else:
        x = natural_lds_sample((init_params, pair_params), node_potentials, num_samples=1, eps=eps, generator=generator)[:, :, 0]
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
                             dense=None, parallel_in_time="auto"):
    """slds_svae.py:159-175, batched: EVERY sequence follows the reference's rule on its own bound -- it stops (its
    factors are frozen) after the first sweep that moved its bound by less than `tol`, so B sequences in one call
    reproduce B reference calls whatever their sweep counts.  The per-mode expected parameters (K small NIW/MNIW
    maps, loop-invariant) are evaluated once per call instead of once per sweep.
    parallel_in_time: "auto" | True | False (use_parallel_in_time)."""
    hmm_global, lds_global = global_natparam
    dense = dense if dense is not None else get_all_lds_local_natparams(lds_global)
    hmm_local = get_hmm_local_natparam(hmm_global)
    pit = use_parallel_in_time(node_potentials[1].shape[0], node_potentials[1].shape[1], parallel_in_time) \
        and not (torch.is_grad_enabled() and any(p.requires_grad for p in node_potentials[:2]))
    lds_stats = initialize_local_meanfield(node_potentials, init_eps, generator, pit)
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
    """slds_svae.py:229-243, summed over the batch; same structure as the global natural parameters."""
    Ez0, Etrans, Ez = hmm_stats
    init_stats, pair_stats = lds_stats
    hmm_global = (Ez0.sum(0), Etrans.sum(0))
    K = Ez.shape[-1]
    lds_global = []
    for k in range(K):
        w0 = Ez[:, 0, k]
        gi = tuple((w0.reshape((-1,) + (1,) * (s.dim() - 1)) * s).sum(0) for s in init_stats)
        wt = Ez[:, 1:, k]
        gp = tuple((wt.reshape(wt.shape + (1,) * (s.dim() - 2)) * s).sum((0, 1)) for s in pair_stats)
        lds_global.append((gaussian.pack_dense(gi[0], gi[1], gi[2], gi[3]), gp))
    return hmm_global, lds_global


def slds_prior_expectedstats(global_natparam):
    """slds_svae.py:272-284."""
    hmm_natparam, lds_natparams = global_natparam
    return hmm_prior_expectedstats(hmm_natparam), [(niw.expectedstats(a), mniw.expectedstats(b)) for a, b in lds_natparams]


def slds_prior_logZ(natparam):
    """slds_svae.py:254-269."""
    (dir_nat, mdir_nat), lds_prior_natparams = natparam
    hmm_term = dirichlet.logZ(dir_nat) + dirichlet.logZ(mdir_nat)
    return hmm_term + sum(niw.logZ(a) + mniw.logZ(b) for a, b in lds_prior_natparams)


def _tuplify(x):
    return tuple(_tuplify(y) for y in x) if isinstance(x, (list, tuple)) else x


def slds_prior_vlb(global_natparam, prior_natparam):
    """slds_svae.py:248-251."""
    expected_stats = slds_prior_expectedstats(global_natparam)
    return contract(sub(_tuplify(prior_natparam), _tuplify(global_natparam)), _tuplify(expected_stats)) \
        - (slds_prior_logZ(prior_natparam) - slds_prior_logZ(global_natparam))


def run_inference(prior_natparam, global_natparam, nn_potentials, num_samples, eps=None, init_eps=None, generator=None):
    """slds_svae.py:289-310: returns (samples, expected_stats, global_vlb, local_vlb)."""
    hmm_global_natparam, lds_global_natparam = global_natparam
    batched = nn_potentials[1].dim() == 3
    pots = tuple(nn_potentials[:2]) if batched else tuple(p[None] for p in nn_potentials[:2])
    dense = get_all_lds_local_natparams(lds_global_natparam)
    with torch.no_grad():
        (hmm_stats, _), (hmm_local_natparam, lds_local_natparam), _ = \
            optimize_local_meanfield(global_natparam, unbox(pots), init_eps=init_eps, generator=generator, dense=dense)
    samples, lds_stats, lds_normalizer = natural_lds_inference_general(lds_local_natparam, pots, num_samples, eps=eps,
                                                                       generator=generator)
    hmm_vlb = get_hmm_vlb(lds_global_natparam, hmm_local_natparam, lds_stats, dense)
    global_lds_stats, local_lds_stats = lds_stats[:-1], lds_stats[-1]
    expected_stats = get_global_stats(hmm_stats, global_lds_stats)
    global_vlb = slds_prior_vlb(global_natparam, prior_natparam)
    lds_vlb = lds_normalizer.sum() - contract(pots, local_lds_stats[:2])
    local_vlb = hmm_vlb.sum() + lds_vlb
    if not batched:
        samples = samples[0]
    return samples, expected_stats, global_vlb, local_vlb
