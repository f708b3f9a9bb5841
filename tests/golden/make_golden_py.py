"""Generates tests/golden/{gmm,iid,slds}_*.npz by RUNNING THE REFERENCE'S OWN PYTHON (staged under Python 3 by
oracle/build_ref_py.py: mechanical py2->3 edits + an autograd->numpy shim, no arithmetic change).

Run in the build container (where /root/reference exists and oracle/build_ref.py has been run):

    python tests/golden/make_golden_py.py

What is executed, reference file:line
  * GMM   svae/models/gmm.py:62-128 local_meanfield (fixed point + evaluation at the optimum), :12-16 run_inference,
          with svae/distributions/{gaussian,niw,dirichlet,categorical}.py.  The random label initialisation
          (gmm.py:126-128) and the sampler's noise (distributions/gaussian.py:27-33) are recorded and injected.
  * iid   svae/models/iid_gausisan.py:42-46 run_inference (noise recorded).
  * SLDS  svae/models/slds_svae.py:80-175, 203-310 -- the module itself cannot be imported (it imports modules that
          are not in the tree), so its functions are exec'd by line range (oracle/build_ref_py.slds_functions) with
          - natural_lds_* = the reference's Cython drivers (lds_inference.py:196-202,232-237,260-264, as slds_svae.py:13-16
            binds them) behind the reference's own canonicalisers (lds_inference.py:62-82),
          - hmm_estep / hmm_logZ = the reference's compiled hmm_logZ (+ hmm_logZ_grad with g = 1: the in-tree
            equivalent of the absent pyhsmm, svae/hmm/hmm_inference.py:65),
          - lds_prior_expectedstats = (unpack_dense(niw.expectedstats), mniw.expectedstats) from svae/distributions
            (what the missing svae.lds.niw / lds_svae supplied: models/lds.py:22-24).
          One sequence per reference call, exactly like the reference.
"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref_py as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
softplus = lambda x: np.log1p(np.exp(x))


def gmm_case(ref, name, K, N, T, S, seed):
    npr = np.random
    npr.seed(seed)
    prior = ref.gmm.init_pgm_param(K, N, alpha=0.05 / K, niw_conc=0.5)
    glob = ref.gmm.init_pgm_param(K, N, alpha=1., niw_conc=1., random_scale=3.)
    node_J = -0.5 * softplus(npr.randn(T, N))
    node_h = 2.0 * npr.randn(T, N)
    label_init = ref.util.normalize(npr.rand(T, K))
    sweeps = [0]
    orig_init, orig_gm = ref.gmm.initialize_meanfield, ref.gmm.gaussian_meanfield

    def counting(*a):
        sweeps[0] += 1
        return orig_gm(*a)
    ref.gmm.initialize_meanfield = lambda lg, pots: label_init
    ref.gmm.gaussian_meanfield = counting
    try:
        (label_stats, gaussian_stats), (dir_stats, niw_stats), (label_nat, gauss_nat), kl = \
            ref.gmm.local_meanfield(glob, (node_J, node_h))
        n_sweeps = sweeps[0] - 1                    # the last call is the evaluation at the optimum (gmm.py:73-75)
        npr.seed(seed + 1)
        eps = npr.randn(T, N, S)                    # what natural_sample will draw (distributions/gaussian.py:29-32)
        npr.seed(seed + 1)
        samples, stats, global_kl, local_kl = ref.gmm.run_inference(prior, glob, (node_J, node_h), S)
    finally:
        ref.gmm.initialize_meanfield, ref.gmm.gaussian_meanfield = orig_init, orig_gm
    # the reference's svae.util.flat is shadowed (util.py:36 binds the name `flatten` that util.py:166 later rebinds to
    # the ravel-concatenate version, so flat(x) == flatten(x)[0] is ONE element): prior_kl (gmm.py:54-58) as checked in
    # therefore multiplies first elements only.  Recorded as-is, plus the intended full inner product.
    full = lambda x: ref.util.flatten(x)
    es = (ref.dirichlet.expectedstats(glob[0]), ref.niw.expectedstats(glob[1]))
    global_kl_full = float(np.dot(full(glob) - full(prior), full(es))
                           - (ref.gmm.prior_logZ(glob) - ref.gmm.prior_logZ(prior)))
    np.savez(os.path.join(OUT, name + ".npz"),
             prior_dir=prior[0], prior_niw=prior[1], glob_dir=glob[0], glob_niw=glob[1],
             node_J=node_J, node_h=node_h, label_init=label_init, eps=eps, n_sweeps=n_sweeps,
             label_stats=label_stats, gaussian_stats=gaussian_stats, dir_stats=dir_stats, niw_stats=niw_stats,
             label_natparam=label_nat, gaussian_natparam=gauss_nat, kl=kl,
             samples=samples, run_label_stats=stats[0], run_niw_stats=stats[1],
             global_kl_as_checked_in=global_kl, global_kl=global_kl_full, local_kl=local_kl)
    print(name, "sweeps", n_sweeps, "kl", float(kl), "global_kl (checked-in / intended)", float(global_kl), global_kl_full)


def iid_case(ref, name, T, n, S, seed):
    npr = np.random
    npr.seed(seed)
    neghalfJ = -0.5 * softplus(npr.randn(T, n))
    h = npr.randn(T, n)
    npr.seed(seed + 1)
    eps = npr.normal(size=(T, S, n))
    npr.seed(seed + 1)
    samples, stats, gvlb, lvlb = ref.iid.run_inference(None, None, (neghalfJ, h, None), S)
    assert stats == () and gvlb == 0.
    np.savez(os.path.join(OUT, name + ".npz"), neghalfJ=neghalfJ, h=h, eps=eps, samples=samples, local_vlb=lvlb)
    print(name, "local_vlb", float(lvlb))


def slds_namespace(ref):
    """The reference's SLDS functions bound as the module docstring says."""
    import numpy as np
    ns = ref.slds._ns
    cy = ref.cython_hmm
    C = lambda x: np.require(x, np.double, 'C')

    def hmm_estep(natparam):
        logZ, saved = cy.hmm_logZ(tuple(C(p) for p in natparam))
        return logZ, tuple(np.asarray(x) for x in cy.hmm_logZ_grad(1.0, saved))
    ns.update(hmm_estep=hmm_estep, hmm_logZ=lambda natparam: cy.hmm_logZ(tuple(C(p) for p in natparam))[0],
              dirichlet=ref.dirichlet, niw=ref.niw, mniw=ref.mniw,
              lds_prior_expectedstats=lambda natparam: (ref.gaussian.unpack_dense(ref.niw.expectedstats(natparam[0])),
                                                        ref.mniw.expectedstats(natparam[1])))
    R._exec_ranges(ns, ref.slds_source, [(118, 128), (203, 226), (248, 251), (254, 269), (272, 284), (289, 310)])
    return ns


def make_slds_global(ref, K, n, rng):
    """Random but well-conditioned global natural parameters in the reference's structure (slds_svae.py:26-75 with the
    argument order of svae/distributions/niw.py:39-42, mniw.py:25-31)."""
    dir_nat = 1.0 + rng.random(K)
    mdir_nat = 1.0 + rng.random((K, K)) + 2.0 * np.eye(K)
    lds = []
    for _ in range(K):
        nu, S, mu, kappa = n + 4. + rng.random(), (n + rng.random()) * np.eye(n), rng.standard_normal(n), 1.0 + rng.random()
        niw_nat = ref.niw.standard_to_natural(S, mu, kappa, nu)
        A = rng.standard_normal((n, n)); A = 0.9 * A / np.abs(np.linalg.eigvals(A)).max()
        nu2, S2, M, Kmat = n + 4. + rng.random(), (n + rng.random()) * np.eye(n), A, (n + rng.random()) * np.eye(n)
        lds.append((niw_nat, ref.mniw.standard_to_natural(nu2, S2, M, Kmat)))
    return (dir_nat, mdir_nat), lds


def slds_case(ref, name, K, n, T, B, S, seed):
    ns = slds_namespace(ref)
    rng = np.random.default_rng(seed)
    glob = make_slds_global(ref, K, n, rng)
    prior = make_slds_global(ref, K, n, rng)
    node_J = -0.5 * softplus(rng.standard_normal((B, T, n)))
    node_h = rng.standard_normal((B, T, n))
    rec = {k: [] for k in ("init_eps", "eps", "samples", "local_vlb", "n_sweeps", "E0", "Etrans", "Ez", "hmm_vlb", "lds_vlb",
                           "g_init_xx", "g_init_x", "g_init_a", "g_init_b", "g_pair_xx", "g_pair_xxn", "g_pair_xnxn",
                           "g_pair_n", "g_E0", "g_Etrans", "init_xx", "init_x", "pair_xx", "pair_xxn", "pair_xnxn",
                           "node_pot_hmm", "lds_lognorm", "E_node_xx", "E_node_x")}
    global_vlb = None
    for b in range(B):
        # a 3-tuple: contract(nn_potentials, local_lds_stats) (slds_svae.py:307) asserts equal structures and the
        # node statistics are (diag ExxT, Ex, 1)
        pots = (node_J[b], node_h[b], np.zeros(T))
        np.random.seed(seed * 1000 + b)
        init_draws = np.random.normal(size=(T, 1, n))      # initialize_local_meanfield's sample path (slds_svae.py:223)
        draws = np.random.normal(size=(T, S, n))           # the final natural_lds_inference_general (slds_svae.py:298)
        sweeps = [0]
        orig = ns["hmm_meanfield"]

        def counting(*a, _o=orig):
            sweeps[0] += 1
            return _o(*a)
        ns["hmm_meanfield"] = counting
        np.random.seed(seed * 1000 + b)
        samples, expected_stats, gvlb, local_vlb = ns["run_inference"](prior, glob, pots, S)
        ns["hmm_meanfield"] = orig
        global_vlb = gvlb
        # the optimum again, piece by piece (deterministic given the same initial draw)
        np.random.seed(seed * 1000 + b)
        (hmm_stats, lds_stats), (hmm_nat, lds_nat), (hmm_vlb, lds_vlb) = ns["optimize_local_meanfield"](glob, pots)
        (g_hmm, g_lds) = expected_stats
        g_lds = list(g_lds)
        # the python sampler consumes the stream backwards in time (lds_inference.py:109-124): re-index by time
        rec["init_eps"].append(init_draws[::-1].copy()); rec["eps"].append(draws[::-1].copy())
        rec["samples"].append(samples); rec["local_vlb"].append(local_vlb); rec["n_sweeps"].append(sweeps[0])
        rec["E0"].append(hmm_stats[0]); rec["Etrans"].append(hmm_stats[1]); rec["Ez"].append(hmm_stats[2])
        rec["hmm_vlb"].append(hmm_vlb); rec["lds_vlb"].append(lds_vlb)
        rec["g_E0"].append(g_hmm[0]); rec["g_Etrans"].append(g_hmm[1])
        for j, key in enumerate(("g_init_xx", "g_init_x", "g_init_a", "g_init_b")):
            rec[key].append(np.stack([np.asarray(g_lds[k][0][j], dtype=float) for k in range(K)]))
        for j, key in enumerate(("g_pair_xx", "g_pair_xxn", "g_pair_xnxn", "g_pair_n")):
            rec[key].append(np.stack([np.asarray(g_lds[k][1][j], dtype=float) for k in range(K)]))
        # statistics of the mean-field optimum (inputs of get_arhmm_local_nodeparams) and its output
        init_stats, pair_stats = lds_stats[0], lds_stats[1]
        rec["init_xx"].append(init_stats[0]); rec["init_x"].append(init_stats[1])
        rec["pair_xx"].append(np.asarray(pair_stats[0])); rec["pair_xxn"].append(np.asarray(pair_stats[1]))
        rec["pair_xnxn"].append(np.asarray(pair_stats[2]))
        rec["node_pot_hmm"].append(ns["get_arhmm_local_nodeparams"](glob[1], lds_stats))
        # the differentiable tail evaluated separately (slds_svae.py:297-300)
        np.random.seed(seed * 1000 + b)
        np.random.normal(size=(T, 1, n))
        _, st, lognorm = ns["natural_lds_inference_general"](lds_nat, pots, S)
        rec["lds_lognorm"].append(lognorm); rec["E_node_xx"].append(st[2][0]); rec["E_node_x"].append(st[2][1])
    out = {k: np.stack([np.asarray(x, dtype=float) for x in v]) for k, v in rec.items()}
    (dir_nat, mdir_nat), lds = glob
    (pdir, pmdir), plds = prior
    out.update(glob_dir=dir_nat, glob_mdir=mdir_nat, glob_niw=np.stack([a for a, _ in lds]),
               prior_dir=pdir, prior_mdir=pmdir, prior_niw=np.stack([a for a, _ in plds]),
               node_J=node_J, node_h=node_h, global_vlb=global_vlb)
    for j, key in enumerate(("A", "B", "C", "d")):
        out["glob_mniw_" + key] = np.stack([np.asarray(bb[j], dtype=float) for _, bb in lds])
        out["prior_mniw_" + key] = np.stack([np.asarray(bb[j], dtype=float) for _, bb in plds])
    np.savez(os.path.join(OUT, name + ".npz"), **out)
    print(name, "sweeps", out["n_sweeps"], "local_vlb", out["local_vlb"], "global_vlb", float(global_vlb))


def main():
    with R.staged() as ref:
        gmm_case(ref, "gmm_K10_N2_T50", 10, 2, 50, 3, 21)
        gmm_case(ref, "gmm_K15_N2_T50", 15, 2, 50, 1, 22)       # the script's own K (experiments/gmm_svae_synth.py:31)
        gmm_case(ref, "gmm_K4_N5_T33", 4, 5, 33, 2, 23)
        iid_case(ref, "iid_T40_n3", 40, 3, 4, 31)
        slds_case(ref, "slds_K3_n3_T12", 3, 3, 12, 3, 2, 41)
        slds_case(ref, "slds_K8_n10_T200", 8, 10, 200, 4, 1, 42)   # BASELINE.json configs[4] shape, B = 4 sequences


if __name__ == "__main__":
    main()
