"""CPU: pins the NumPy oracle (oracle/lds_numpy.py) to the reference.

(1) against the committed golden vectors, which tests/golden/make_golden.py produced by running the
    reference's own compiled Cython kernels (oracle/_ref) in the build container;
(2) live against oracle/_ref when it is present (build container and GPU box both have the .so);
(3) through the exponential-family identity d lognorm / d(node natparams) = E_node stats, the property
    the reference authors relied on (lds_inference.py:205-218).
"""
import numpy as np
import pytest
from oracle import lds_numpy as O
from oracle import ref
from util import golden_cases, load, rel_err, golden_natparams

TOL = 1e-9   # oracle vs reference: both fp64 on the same algorithm


@pytest.mark.parametrize("path", golden_cases("lds_"))
def test_oracle_matches_golden(path):
    d = load(path)
    init, pair, node = golden_natparams(d)
    samples, (E_init, E_pair, E_node), lognorm = O.natural_lds_inference_general((init, pair), node, d["eps"])
    assert rel_err(samples, d["samples"]) < TOL
    assert rel_err(lognorm, d["lognorm"]) < TOL
    assert rel_err(E_init[0], d["E_init_xx"]) < TOL and rel_err(E_init[1], d["E_init_x"]) < TOL
    assert rel_err(E_pair[0], d["E_pair_xx"]) < TOL
    assert rel_err(E_pair[1], d["E_pair_xxn"]) < TOL
    assert rel_err(E_pair[2], d["E_pair_xnxn"]) < TOL
    assert np.array_equal(E_pair[3], d["E_pair_n"])
    assert rel_err(E_node[0], d["E_node_xx"]) < TOL and rel_err(E_node[1], d["E_node_x"]) < TOL
    (Jp, hp), (Jf, hf) = O.natural_filter_forward_general(init, pair, node)[0]
    assert rel_err(Jp, d["J_pred"]) < TOL and rel_err(hp, d["h_pred"]) < TOL
    assert rel_err(Jf, d["J_filt"]) < TOL and rel_err(hf, d["h_filt"]) < TOL


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,T,S", [(4, 9, 2), (10, 30, 1)])
def test_oracle_matches_compiled_reference_live(n, T, S):
    rng = np.random.default_rng(n * 100 + T)
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(3, T, n, rng)
    outs = [ref.lds_estep_one(init, pair, tuple(x[b] for x in node), S, seed=b) for b in range(3)]
    eps = np.stack([o["eps"] for o in outs])
    samples, (E_init, E_pair, E_node), lognorm = O.natural_lds_inference_general((init, pair), node, eps)
    for b, o in enumerate(outs):
        assert rel_err(samples[b], o["samples"]) < TOL
        assert abs(lognorm[b] - o["lognorm"]) < TOL * abs(o["lognorm"])
        for k in range(3):
            assert rel_err(E_pair[k][b], o["E_pair"][k]) < TOL
        assert rel_err(E_node[0][b], o["E_node"][0]) < TOL and rel_err(E_node[1][b], o["E_node"][1]) < TOL


def test_expfam_identity_finite_difference():
    """d lognorm/d h_node[t,i] = E[x_t]_i and d lognorm/d J_node[t,i] = E[x_t x_t']_ii."""
    rng = np.random.default_rng(5)
    n, T = 3, 6
    init, pair = O.rand_lds_natparam(n, rng)
    J, h, logZ = O.rand_node_potentials(1, T, n, rng)
    lognorm, (_, _, E_node) = O.natural_lds_estep_general((init, pair), (J, h, logZ))
    step = 1e-6
    for (t, i) in [(0, 0), (2, 1), (5, 2)]:
        hp, hm = h.copy(), h.copy(); hp[0, t, i] += step; hm[0, t, i] -= step
        fd = (O.natural_lds_estep_general((init, pair), (J, hp, logZ))[0]
              - O.natural_lds_estep_general((init, pair), (J, hm, logZ))[0]) / (2 * step)
        assert abs(fd[0] - E_node[1][0, t, i]) < 1e-6 * max(1.0, abs(fd[0]))
        Jp, Jm = J.copy(), J.copy(); Jp[0, t, i] += step; Jm[0, t, i] -= step
        fd = (O.natural_lds_estep_general((init, pair), (Jp, h, logZ))[0]
              - O.natural_lds_estep_general((init, pair), (Jm, h, logZ))[0]) / (2 * step)
        assert abs(fd[0] - E_node[0][0, t, i]) < 1e-6 * max(1.0, abs(fd[0]))


def test_inhomogeneous_pair_params_reduce_to_homogeneous():
    rng = np.random.default_rng(8)
    n, T, B = 3, 5, 2
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    eps = rng.standard_normal((B, T, 1, n))
    rep = tuple(np.broadcast_to(x, (T - 1,) + np.shape(x)).copy() for x in pair)
    s0, (Ei0, Ep0, En0), l0 = O.natural_lds_inference_general((init, pair), node, eps)
    s1, (Ei1, Ep1, En1), l1 = O.natural_lds_inference_general((init, rep), node, eps)
    assert rel_err(s1, s0) < 1e-12 and rel_err(l1, l0) < 1e-12
    assert Ep1[0].shape == (B, T - 1, n, n)
    for k in range(3):
        assert rel_err(Ep1[k].sum(1), Ep0[k]) < 1e-12
