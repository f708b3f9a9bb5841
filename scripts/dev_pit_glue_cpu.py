"""Dev: run svae_b200/lds/lds_pit.py's glue on CPU with the two kernel entry points replaced by oracle-backed shims,
to check indexing/shapes before spending GPU time.  Not part of the product or the test-suite."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
from oracle import lds_numpy as O
import pit_numpy as P
from svae_b200.lds import lds_pit, lds_inference as K

T64 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)
N = lambda x: None if x is None else x.detach().numpy()
lds_pit._dev = lambda h: torch.device("cpu")


def fake_compose(pair_params, node_params, chunk_len):
    pair = tuple(N(T64(x)) for x in pair_params[:3]) + (N(T64(pair_params[3])),)
    node = tuple(N(T64(x)) for x in node_params)
    if len(node) == 2:
        node = node + (np.zeros(node[1].shape[:2]),)
    cp, e1, e2 = P.compose_host(pair, node, chunk_len)
    return tuple(T64(x) for x in cp), T64(e1), T64(e2)


def fake_forward(init_params, pair_params, node_params, num_samples=1, eps=None, per_t_pair_stats=False, want_stats=True,
                 messages_out=False, term_J=None, **kw):
    nJ, nh = N(T64(node_params[0])), N(T64(node_params[1]))
    nlogZ = N(T64(node_params[2])) if len(node_params) > 2 and node_params[2] is not None else np.zeros(nh.shape[:2])
    B, T, n = nh.shape
    S = int(num_samples or 0)
    iJ, ih = N(T64(init_params[0])), N(T64(init_params[1]))
    ilZ = float(sum(init_params[2:])) if len(init_params) > 2 and iJ.ndim == 2 else 0.
    pair = tuple(N(T64(x)) for x in pair_params)
    out = dict(B=B, T=T, n=n, S=S, batched=True, status=torch.zeros(1, dtype=torch.int32))
    keys = ["lognorm", "E_node_xx", "E_node_x", "samples", "E_init_xx", "E_init_x", "E_pair_xx", "E_pair_xxn", "E_pair_xnxn",
            "Exx_t", "Exxn_t", "J_pred", "h_pred"]
    acc = {k: [] for k in keys}
    for b in range(B):
        init_b = (iJ[b] if iJ.ndim == 3 else iJ, ih[b] if ih.ndim == 2 else ih, ilZ)
        pair_b = tuple(x[b] if (x.ndim in (4,) or (k == 3 and x.ndim == 2)) else x for k, x in enumerate(pair))
        Jd = nJ[b:b + 1, :, :, None] * np.eye(n)
        if term_J is not None:
            Jd = Jd.copy(); Jd[0, -1] += N(T64(term_J))[b]
        node_b = (Jd, nh[b:b + 1], nlogZ[b:b + 1])
        e = N(T64(eps))[b:b + 1] if S else np.zeros((1, T, 0, n))
        msgs, ln = O.natural_filter_forward_general(init_b, pair_b, node_b)
        Ei, Ep, En = O.natural_smoother_general(msgs, pair_b, diagonal_node_stats=False)
        x = O.natural_sample_backward_general(msgs, pair_b, e) if S else None
        homog = pair_b[0].ndim == 2
        acc["lognorm"].append(ln[0]); acc["E_node_xx"].append(np.diagonal(En[0][0], axis1=-1, axis2=-2)); acc["E_node_x"].append(En[1][0])
        acc["Exx_t"].append(En[0][0]); acc["J_pred"].append(msgs[0][0][0]); acc["h_pred"].append(msgs[0][1][0])
        acc["E_init_xx"].append(Ei[0][0]); acc["E_init_x"].append(Ei[1][0])
        if homog:
            for k, key in enumerate(("E_pair_xx", "E_pair_xxn", "E_pair_xnxn")):
                acc[key].append(Ep[k][0])
        else:
            acc["Exxn_t"].append(Ep[1][0])
        if S:
            acc["samples"].append(x[0])
    for k in keys:
        out[k] = T64(np.stack(acc[k])) if acc[k] else None
    return out


lds_pit.compose_chunks = fake_compose
K.lds_estep_forward = fake_forward
K._lib.lib = lambda: None

err = lambda a, b: float(np.abs(N(a) - b).max() / max(np.abs(b).max(), 1e-300))
for (n, T, B, S, Lc) in [(4, 25, 2, 2, 6), (3, 14, 3, 1, 5), (10, 60, 2, 1, None), (16, 30, 1, 1, 9), (8, 20, 2, 0, 7), (2, 90, 1, 3, 20),
                         (14, 22, 1, 4, 5), (15, 22, 1, 15, 5)]:
    rng = np.random.default_rng(50 + n)
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    node = (node[0], node[1], 0.3 * rng.standard_normal((B, T)))
    eps = rng.standard_normal((B, T, S, n))
    w_samples, (wE_init, wE_pair, wE_node), w_lognorm = O.natural_lds_inference_general((init, pair), node, eps)
    g = lambda tree: tuple(T64(x) for x in tree)
    samples, (E_init, E_pair, E_node), lognorm = lds_pit.natural_lds_inference_parallel(
        (g(init), g(pair)), g(node), num_samples=S, eps=T64(eps) if S else None, chunk_len=Lc)
    print((n, T, B, S, Lc), "lognorm %.1e" % err(lognorm, w_lognorm), "Einit %.1e %.1e" % (err(E_init[0], wE_init[0]), err(E_init[1], wE_init[1])),
          "Epair %.1e %.1e %.1e" % tuple(err(E_pair[k], wE_pair[k]) for k in range(3)),
          "Enode %.1e %.1e" % (err(E_node[0], wE_node[0]), err(E_node[1], wE_node[1])),
          "samples %.1e" % (err(samples, w_samples) if S else 0.))


# time-inhomogeneous pair parameters
for per_seq in (False, True):
    n, T, B, S, Lc = 5, 23, 2, 1, 4
    rng = np.random.default_rng(7)
    init, _ = O.rand_lds_natparam(n, rng)
    shape = (B, T - 1) if per_seq else (T - 1,)
    prs = [O.rand_lds_natparam(n, rng)[1] for _ in range(int(np.prod(shape)))]
    pair = tuple(np.stack([p[k] for p in prs]).reshape(shape + np.shape(prs[0][k])) for k in range(4))
    node = O.rand_node_potentials(B, T, n, rng)
    eps = rng.standard_normal((B, T, S, n))
    want = [O.natural_lds_inference_general((init, tuple(x[q] for x in pair) if per_seq else pair), tuple(x[q:q + 1] for x in node), eps[q:q + 1]) for q in range(B)]
    samples, (E_init, E_pair, E_node), lognorm = lds_pit.natural_lds_inference_parallel(
        (g(init), g(pair)), g(node), num_samples=S, eps=T64(eps), chunk_len=Lc)
    cat = lambda f: np.concatenate([f(w) for w in want])
    print("inhomog per_seq=%s" % per_seq, "lognorm %.1e" % err(lognorm, cat(lambda w: w[2])), "samples %.1e" % err(samples, cat(lambda w: w[0])),
          "Epair %.1e %.1e %.1e" % tuple(err(E_pair[k], cat(lambda w: w[1][1][k])) for k in range(3)),
          "Enode %.1e %.1e" % tuple(err(E_node[k], cat(lambda w: w[1][2][k])) for k in range(2)))
