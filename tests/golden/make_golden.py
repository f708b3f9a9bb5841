"""Generates tests/golden/lds_*.npz from the reference's OWN compiled Cython kernels (oracle/_ref).

Run in the build container (where /root/reference exists and oracle/build_ref.py has been run):

    python tests/golden/make_golden.py

Each fixture holds the natural-parameter inputs of one small LDS problem and what the reference's
natural_filter_forward_general -> natural_smoother_general -> natural_sample_backward
(svae/lds/cython_lds_inference.pyx:28-90,149-210,310-355) returned for each of its B sequences, one
reference call per sequence, plus the noise the reference's sampler drew (re-indexed by time).
The VJP fixtures come from natural_sample_backward_grad / natural_smoother_general_grad /
natural_filter_grad (pyx:357-409,236-306,92-145) composed as lds_inference.py:26-39 wires them.
"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref, lds_numpy as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [  # name, n, T, B, S, seed
    ("lds_n2_T5", 2, 5, 3, 2, 11),
    ("lds_n3_T2", 3, 2, 2, 1, 12),
    ("lds_n8_T50", 8, 50, 4, 1, 13),
    ("lds_n10_T200", 10, 200, 2, 1, 14),
    ("lds_n16_T40", 16, 40, 2, 3, 15),
    ("lds_n1_T7", 1, 7, 2, 1, 16),
]


def main():
    for name, n, T, B, S, seed in CASES:
        rng = np.random.default_rng(seed)
        init, pair = O.rand_lds_natparam(n, rng)
        node = O.rand_node_potentials(B, T, n, rng)
        node = (node[0], node[1], 0.1 * rng.standard_normal((B, T)))
        outs = [ref.lds_estep_one(init, pair, tuple(x[b] for x in node), S, seed=seed * 100 + b) for b in range(B)]
        st = lambda f: np.stack([f(o) for o in outs])
        d = dict(
            init_J=init[0], init_h=init[1], init_logZ=init[2],
            J11=pair[0], J12=pair[1], J22=pair[2], pair_logZ=pair[3],
            node_J=node[0], node_h=node[1], node_logZ=node[2],
            eps=st(lambda o: o["eps"]), samples=st(lambda o: o["samples"]), lognorm=st(lambda o: o["lognorm"]),
            E_init_xx=st(lambda o: o["E_init"][0]), E_init_x=st(lambda o: o["E_init"][1]),
            E_pair_xx=st(lambda o: o["E_pair"][0]), E_pair_xxn=st(lambda o: o["E_pair"][1]),
            E_pair_xnxn=st(lambda o: o["E_pair"][2]), E_pair_n=st(lambda o: o["E_pair"][3]),
            E_node_xx=st(lambda o: o["E_node"][0]), E_node_x=st(lambda o: o["E_node"][1]),
            J_pred=st(lambda o: o["messages"][0][0]), h_pred=st(lambda o: o["messages"][0][1]),
            J_filt=st(lambda o: o["messages"][1][0]), h_filt=st(lambda o: o["messages"][1][1]),
        )
        # VJP fixture: random cotangents on everything the E-step returns
        g_samples = rng.standard_normal((B, T, S, n))
        g_Einit = (rng.standard_normal((B, n, n)), rng.standard_normal((B, n)))
        g_Epair = tuple(rng.standard_normal((B, n, n)) for _ in range(3))
        g_Enode = (rng.standard_normal((B, T, n)), rng.standard_normal((B, T, n)))
        g_lognorm = rng.standard_normal(B)
        vj = []
        for b in range(B):
            gs = ((g_Einit[0][b], g_Einit[1][b], 0., 0.),
                  (g_Epair[0][b], g_Epair[1][b], g_Epair[2][b], 0.),
                  (g_Enode[0][b].copy(), g_Enode[1][b].copy(), np.zeros(T)))
            vj.append(ref.lds_vjp_one(init, pair, tuple(x[b] for x in node), g_samples[b], gs, g_lognorm[b],
                                      seed=seed * 100 + b, num_samples=S))
            assert np.allclose(vj[-1]["eps"], d["eps"][b])
        d.update(g_samples=g_samples, g_E_init_xx=g_Einit[0], g_E_init_x=g_Einit[1],
                 g_E_pair_xx=g_Epair[0], g_E_pair_xxn=g_Epair[1], g_E_pair_xnxn=g_Epair[2],
                 g_E_node_xx=g_Enode[0], g_E_node_x=g_Enode[1], g_lognorm=g_lognorm,
                 vjp_node_J=np.stack([v["g_J_node"] for v in vj]), vjp_node_h=np.stack([v["g_h_node"] for v in vj]),
                 vjp_node_logZ=np.stack([v["g_logZ_node"] for v in vj]))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "ok", {k: np.shape(v) for k, v in list(d.items())[:3]})


if __name__ == "__main__":
    main()
