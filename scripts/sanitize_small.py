"""Dev helper for compute-sanitizer: one small launch of every kernel and mode (general + packed forward, SAVE, VJP with its
message ports, masked SLDS sweep with the glue kernels, HMM, Viterbi, GMM fixed point, reductions)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from svae_b200 import synthetic as SY
from svae_b200.lds import lds_inference as K
from svae_b200.hmm import hmm_inference as HK
from svae_b200.models import slds_svae as SL, gmm as GM
from svae_b200.distributions import dirichlet as DD, niw as NW
rng = np.random.default_rng(0)
g = lambda x, grad=False: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda").requires_grad_(grad)
for n, T, B, S in [(10, 12, 5, 1), (16, 6, 3, 2), (3, 5, 7, 1), (7, 4, 9, 0), (8, 3, 6, 3), (10, 1, 4, 1), (11, 2, 3, 2)]:
    init, pair = SY.lds_natparam(n, rng)
    node = SY.node_potentials(B, T, n, rng)[:2]
    gi, gp = tuple(g(x) for x in init), tuple(g(x) for x in pair)
    nJ, nh = g(node[0], True), g(node[1], True)
    samples, (Ei, Ep, En), ln = K.natural_lds_inference_general((gi, gp), (nJ, nh), S)        # packed SAVE (S <= 2) or general + VJP
    ((samples.sum() if S else 0.) + ln.sum() + En[0].sum() + Ep[1].sum()).backward()
    o = K.lds_estep_forward(gi, gp, (nJ.detach(), nh.detach()), S)                             # packed / general, no save
    K.lds_reduce_global_stats(o)
    # the three primitives, differentiated one by one (message ports of both kernels)
    nJ2, nh2 = g(node[0], True), g(node[1], True)
    msgs, ln2 = K.natural_filter_forward_general(gi, gp, (nJ2, nh2))
    st = K.natural_smoother_general(msgs, gp)
    smp = K.natural_sample_backward(msgs, gp, max(S, 1))
    (ln2.sum() + st[2][1].sum() + st[1][1].sum() + smp.sum()).backward()
lp, lA = SY.hmm_natparam(7, rng)
nat = (g(lp), g(lA), g(rng.standard_normal((5, 9, 7))))
HK.hmm_estep(nat); HK.hmm_viterbi(nat)
# SLDS: masked sweeps on persistent buffers (glue kernels, masked HMM / LDS launches) + differentiable tail
Ks, ns, Ts, Bs = 3, 4, 9, 5
(dn, mn), lds_np = SY.slds_global_natparam(Ks, ns, rng)
glob = ((g(dn), g(mn)), [(g(a), tuple(g(x) for x in b)) for a, b in lds_np])
sJ, sh = g(SY.node_potentials(Bs, Ts, ns, rng)[0], True), g(SY.node_potentials(Bs, Ts, ns, rng)[1], True)
s_, st_, gv, lv = SL.run_inference(glob, glob, (sJ, sh), 2)
(lv + s_.sum()).backward()
# GMM fixed point
dg, ng = SY.gmm_global_natparam(4, 2, rng)
li = torch.rand(3, 20, 4, dtype=torch.float64, device="cuda"); li = li / li.sum(-1, keepdim=True)
GM.meanfield_fixed_point_cuda(DD.expectedstats(g(dg)), NW.expectedstats(g(ng)), -0.5 * torch.rand(3, 20, 2, dtype=torch.float64, device="cuda") - 0.1,
                              torch.randn(3, 20, 2, dtype=torch.float64, device="cuda"), li)
torch.cuda.synchronize()
print("sanitize run ok")
