"""Dev: numerical stress of the parallel-in-time pipeline on CPU (oracle-backed shims): large posterior means (the
boundary covariance is formed as E[xx'] - mu mu'), long chunks with strongly contractive gains, weak node potentials."""
import sys, os, runpy
sys.argv = [sys.argv[0]]
ns = runpy.run_path(os.path.join(os.path.dirname(__file__), "dev_pit_glue_cpu.py"), run_name="shim")
import numpy as np, torch
O, lds_pit, T64, err = ns["O"], ns["lds_pit"], ns["T64"], ns["err"]
g = lambda tree: tuple(T64(x) for x in tree)
for name, hscale, jscale, T, Lc in [("means ~1e3", 1e3, 1.0, 60, 8), ("means ~1e5", 1e5, 1.0, 60, 8), ("weak nodes", 1.0, 1e-6, 80, 10),
                                    ("long chunks", 1.0, 1.0, 400, 100)]:
    rng = np.random.default_rng(1)
    n, B, S = 6, 2, 1
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    node = (node[0] * jscale, node[1] * hscale, node[2])
    eps = rng.standard_normal((B, T, S, n))
    w_samples, (wE_init, wE_pair, wE_node), w_lognorm = O.natural_lds_inference_general((init, pair), node, eps)
    samples, (E_init, E_pair, E_node), lognorm = lds_pit.natural_lds_inference_parallel((g(init), g(pair)), g(node), S, T64(eps), chunk_len=Lc)
    print("%-12s |mu|max %.1e  lognorm %.1e  Enode %.1e %.1e  Epair %.1e  samples %.1e" % (
        name, np.abs(wE_node[1]).max(), err(lognorm, w_lognorm), err(E_node[0], wE_node[0]), err(E_node[1], wE_node[1]),
        max(err(E_pair[k], wE_pair[k]) for k in range(3)), err(samples, w_samples)))
