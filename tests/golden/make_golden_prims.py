"""Generates tests/golden/ldsprim_*.npz: the RAW outputs of the reference's three primitive VJPs (oracle/_ref),
one sequence per call:

    natural_filter_grad((g_messages, g_lognorm), saved)       cython_lds_inference.pyx:92-145
    natural_smoother_general_grad(g_stats, saved)             pyx:236-306  -> cotangents of the forward messages
    natural_sample_backward_grad(g_samples, saved)            pyx:357-409  -> cotangents of the forward messages

    python tests/golden/make_golden_prims.py
"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref, lds_numpy as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [("ldsprim_n3_T5", 3, 5, 3, 2, 51), ("ldsprim_n10_T30", 10, 30, 2, 1, 52), ("ldsprim_n16_T8", 16, 8, 2, 1, 53)]


def main():
    lds, _ = ref.load()
    # always COPIES: the reference's *_grad functions accumulate into their cotangent arguments in place
    # (np.require(g.T, 'F') aliases a C-ordered (T,n) input, cython_lds_inference.pyx:95-98)
    c = lambda x: np.array(x, dtype=np.float64, order='C', copy=True)
    for name, n, T, B, S, seed in CASES:
        rng = np.random.default_rng(seed)
        init, pair = O.rand_lds_natparam(n, rng)
        node = O.rand_node_potentials(B, T, n, rng)
        node = (node[0], node[1], 0.1 * rng.standard_normal((B, T)))
        sym = lambda a: a + np.swapaxes(a, -1, -2)
        g_msg = dict(gJp=sym(rng.standard_normal((B, T, n, n))), ghp=rng.standard_normal((B, T, n)),
                     gJf=sym(rng.standard_normal((B, T, n, n))), ghf=rng.standard_normal((B, T, n)), gln=rng.standard_normal(B))
        g_samples = rng.standard_normal((B, T, S, n))
        g_Einit = (sym(rng.standard_normal((B, n, n))), rng.standard_normal((B, n)))
        g_Epair = (sym(rng.standard_normal((B, n, n))), rng.standard_normal((B, n, n)), sym(rng.standard_normal((B, n, n))))
        g_Enode = (rng.standard_normal((B, T, n)), rng.standard_normal((B, T, n)))
        rec = {k: [] for k in ("filt_gJ", "filt_gh", "filt_glogZ", "sm_gJp", "sm_ghp", "sm_gJf", "sm_ghf",
                               "sa_gJp", "sa_ghp", "sa_gJf", "sa_ghf", "eps", "J_pred", "h_pred", "J_filt", "h_filt")}
        for b in range(B):
            nd = tuple(c(x[b]) for x in node)
            (msgs, lognorm), saved_f = lds.natural_filter_forward_general(tuple(init), tuple(pair), nd)
            gm = ((c(g_msg["gJp"][b]), c(g_msg["ghp"][b])), (c(g_msg["gJf"][b]), c(g_msg["ghf"][b])))
            gn = lds.natural_filter_grad((gm, float(g_msg["gln"][b])), saved_f)
            rec["filt_gJ"].append(np.asarray(gn[0])); rec["filt_gh"].append(np.asarray(gn[1])); rec["filt_glogZ"].append(np.asarray(gn[2]))
            # the filter's saved intermediates are consumed by its grad: run the primitives again for the other two
            (msgs, lognorm), _ = lds.natural_filter_forward_general(tuple(init), tuple(pair), nd)
            stats, saved_s = lds.natural_smoother_general(msgs, tuple(pair))
            gst = ((c(g_Einit[0][b]), c(g_Einit[1][b]), 0., 0.), (c(g_Epair[0][b]), c(g_Epair[1][b]), c(g_Epair[2][b]), 0.),
                   (c(g_Enode[0][b]), c(g_Enode[1][b]), np.zeros(T)))
            gsm = lds.natural_smoother_general_grad(gst, saved_s)
            np.random.seed(seed * 10 + b)
            samples, saved_b = lds.natural_sample_backward(msgs, tuple(pair), S)
            rec["eps"].append(np.ascontiguousarray(np.asarray(saved_b[3]).T.copy()))
            gsa = lds.natural_sample_backward_grad(np.array(g_samples[b], dtype=np.float64, order="C"), saved_b)
            for key, val in (("sm_gJp", gsm[0][0]), ("sm_ghp", gsm[0][1]), ("sm_gJf", gsm[1][0]), ("sm_ghf", gsm[1][1]),
                             ("sa_gJp", gsa[0][0]), ("sa_ghp", gsa[0][1]), ("sa_gJf", gsa[1][0]), ("sa_ghf", gsa[1][1]),
                             ("J_pred", msgs[0][0]), ("h_pred", msgs[0][1]), ("J_filt", msgs[1][0]), ("h_filt", msgs[1][1])):
                rec[key].append(np.array(val, dtype=np.float64))
        d = {k: np.stack(v) for k, v in rec.items()}
        d.update(init_J=init[0], init_h=init[1], init_logZ=init[2], J11=pair[0], J12=pair[1], J22=pair[2], pair_logZ=pair[3],
                 node_J=node[0], node_h=node[1], node_logZ=node[2], g_samples=g_samples,
                 g_E_init_xx=g_Einit[0], g_E_init_x=g_Einit[1], g_E_pair_xx=g_Epair[0], g_E_pair_xxn=g_Epair[1],
                 g_E_pair_xnxn=g_Epair[2], g_E_node_xx=g_Enode[0], g_E_node_x=g_Enode[1], **{"in_" + k: v for k, v in g_msg.items()})
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "ok")


if __name__ == "__main__":
    main()
