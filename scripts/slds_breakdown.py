import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svae_b200.models import slds_svae as M
from svae_b200.lds.lds_inference import natural_lds_estep_general
from svae_b200.hmm.hmm_inference import hmm_estep
K, n, T, B = 8, 10, 200, 1024
g = torch.Generator().manual_seed(0)
glob = M.make_slds_global_natparam(K, n, alpha=2., sticky_bias=1., random=True, generator=g, device="cuda")
nJ = -0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64, device="cuda"))
nh = torch.randn(B, T, n, dtype=torch.float64, device="cuda")
hmm_global, lds_global = glob
def timed(f, *a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(*a); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
with torch.no_grad():
    lds_stats = M.initialize_local_meanfield((nJ, nh))
    dense = M.get_all_lds_local_natparams(lds_global)
    for rep in range(3):
        node, t1 = timed(M.get_arhmm_local_nodeparams, lds_global, lds_stats, dense)
        ip, pp = M.get_hmm_local_natparam(hmm_global)
        (vl, hs), t2 = timed(hmm_estep, (ip, pp, node))
        nat, t3 = timed(M.get_var_lds_local_natparam, lds_global, hs, dense)
        (lv, lds_stats), t4 = timed(natural_lds_estep_general, nat, (nJ, nh))
        print("arhmm node potentials %.2f ms | hmm_estep %.2f ms | mix pair params %.2f ms | lds estep (inhomog, per-t stats) %.2f ms" % (t1, t2, t3, t4))
