"""Dev timing: sequential fused kernel vs parallel-in-time driver in the small-B / long-T regime."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from oracle import lds_numpy as O
from svae_b200.lds import lds_inference as K, lds_pit


def timeit(f, it=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")
for n, T, B, S in [(10, 20000, 1, 1), (10, 2000, 16, 1), (10, 100000, 1, 1), (10, 5000, 4, 0), (16, 4000, 8, 1), (10, 200, 64, 1)]:
    rng = np.random.default_rng(0)
    init, pair = O.rand_lds_natparam(n, rng)
    node = O.rand_node_potentials(B, T, n, rng)
    nat = (tuple(g(x) for x in init), tuple(g(x) for x in pair))
    nd = tuple(g(x) for x in node)
    eps = torch.randn((B, T, S, n), dtype=torch.float64, device="cuda") if S else None
    t_seq = timeit(lambda: K.natural_lds_inference_general(nat, nd, S, eps))
    row = "n=%d T=%d B=%d S=%d  sequential %.2f ms |" % (n, T, B, S, t_seq)
    for Lc in (None, 32, 64, 128, 256):
        if Lc and Lc >= T - 1:
            continue
        t_par = timeit(lambda: lds_pit.natural_lds_inference_parallel(nat, nd, S, eps, chunk_len=Lc))
        row += " Lc=%s %.2f ms (x%.1f)" % (Lc or "sqrt", t_par, t_seq / t_par)
    print(row, flush=True)

from oracle import hmm_numpy as HO
from svae_b200.hmm import hmm_inference as H, hmm_pit
for K, T, B in [(32, 100000, 1), (32, 20000, 4), (8, 100000, 1), (32, 2000, 16), (32, 500, 64)]:
    rng = np.random.default_rng(0)
    lp, lA = HO.rand_hmm_natparam(K, rng)
    nat = (g(lp), g(lA), torch.randn((B, T, K), dtype=torch.float64, device="cuda"))
    t_seq = timeit(lambda: H.hmm_estep(nat))
    row = "HMM K=%d T=%d B=%d  sequential %.2f ms |" % (K, T, B, t_seq)
    for Lc in (None, 64, 256, 1024):
        if Lc and Lc >= T - 1:
            continue
        t_par = timeit(lambda: hmm_pit.hmm_estep_parallel(nat, chunk_len=Lc))
        row += " Lc=%s %.2f ms (x%.1f)" % (Lc or "sqrt", t_par, t_seq / t_par)
    print(row, flush=True)

# stage breakdown
n, T, B = 10, 100000, 1
rng = np.random.default_rng(0)
init, pair = O.rand_lds_natparam(n, rng)
node = O.rand_node_potentials(B, T, n, rng)
gp, nd = tuple(g(x) for x in pair), tuple(g(x) for x in node)
for Lc in (128, 316):
    print("LDS compose alone T=%d Lc=%d: %.2f ms" % (T, Lc, timeit(lambda: lds_pit.compose_chunks(gp, nd, Lc))))
K = 32
lp, lA = HO.rand_hmm_natparam(K, rng)
hn = torch.randn((1, T, K), dtype=torch.float64, device="cuda")
L_ = H._lib.lib()
for Lc in (316, 1024):
    C = L_.svae_lds_compose_num_chunks(T, Lc)
    P_, ls = torch.empty((1, C, K, K), dtype=torch.float64, device="cuda"), torch.empty((1, C), dtype=torch.float64, device="cuda")
    f = lambda: L_.svae_hmm_compose_chunks(1, T, K, Lc, H._lib.ptr(g(lA)), H._lib.ptr(hn), H._lib.ptr(P_), H._lib.ptr(ls), H._lib.stream_ptr(hn.device))
    print("HMM compose alone T=%d Lc=%d: %.2f ms" % (T, Lc, timeit(f)))

# SLDS local mean field on one long sequence
import time
from svae_b200.models import slds_svae as M
Ks, ns, Ts, Bs = 8, 10, 20000, 1
gg = torch.Generator().manual_seed(0)
glob = M.make_slds_global_natparam(Ks, ns, alpha=2., sticky_bias=1., random=True, generator=gg, device="cuda")
nJ = -0.5 * torch.nn.functional.softplus(torch.randn(Bs, Ts, ns, dtype=torch.float64, device="cuda"))
nh = torch.randn(Bs, Ts, ns, dtype=torch.float64, device="cuda")
ie = torch.randn(Bs, Ts, 1, ns, dtype=torch.float64, device="cuda")
for flag in (False, True, False, True):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        out = M.optimize_local_meanfield(glob, (nJ, nh), init_eps=ie, parallel_in_time=flag)
    torch.cuda.synchronize()
    print("SLDS K=%d d=%d T=%d B=%d optimize_local_meanfield parallel_in_time=%s: %.1f ms (vlb %.6f)" %
          (Ks, ns, Ts, Bs, flag, (time.perf_counter() - t0) * 1e3, float((out[2][0] + out[2][1]).sum())), flush=True)
