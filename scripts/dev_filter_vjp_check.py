"""Dev: which message-cotangent port of the filter VJP disagrees with torch autograd through a CPU filter."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import lds_torch as OT
from util import load, golden_natparams
from svae_b200.lds import lds_inference as K
d = load("tests/golden/ldsprim_n3_T5.npz")
init, pair, node = golden_natparams(d)
t64 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)
g = lambda x: t64(x).cuda()
B, T, n = d["node_h"].shape
def cpu_filter(nJ, nh, nZ):
    J, h = t64(init[0]).expand(B, n, n), t64(init[1]).expand(B, n)
    ln = torch.zeros(B, dtype=torch.float64) + float(init[2])
    Jp, hp, Jf, hf = [], [], [], []
    eye = torch.eye(n, dtype=torch.float64)
    pr = tuple(t64(x) for x in pair)
    for t in range(T):
        Jp.append(J); hp.append(h)
        J, h = J + nJ[:, t, :, None] * eye, h + nh[:, t]
        ln = ln + nZ[:, t]
        Jf.append(J); hf.append(h)
        if t < T - 1:
            (J, h), term = OT.natural_predict(J, h, *pr)
            ln = ln + term
    ln = ln + OT.natural_lognorm(J, h)
    return torch.stack(Jp, 1), torch.stack(hp, 1), torch.stack(Jf, 1), torch.stack(hf, 1), ln
for which in (["gln"], ["gJf"], ["ghf"], ["gJp"], ["ghp"], ["gJp", "ghp", "gJf", "ghf", "gln"]):
    w = {k: (t64(d["in_" + k]) if k in which else torch.zeros_like(t64(d["in_" + k]))) for k in ("gJp", "ghp", "gJf", "ghf", "gln")}
    nJ, nh, nZ = (t64(d[k]).requires_grad_(True) for k in ("node_J", "node_h", "node_logZ"))
    Jp, hp, Jf, hf, ln = cpu_filter(nJ, nh, nZ)
    ((Jp * w["gJp"]).sum() + (hp * w["ghp"]).sum() + (Jf * w["gJf"]).sum() + (hf * w["ghf"]).sum() + (ln * w["gln"]).sum()).backward()
    gJ, gh, gZ = (g(d[k]).requires_grad_(True) for k in ("node_J", "node_h", "node_logZ"))
    ((Jp2, hp2), (Jf2, hf2)), ln2 = K.natural_filter_forward_general(tuple(g(x) for x in init), tuple(g(x) for x in pair), (gJ, gh, gZ))
    ((Jp2 * w["gJp"].cuda()).sum() + (hp2 * w["ghp"].cuda()).sum() + (Jf2 * w["gJf"].cuda()).sum() + (hf2 * w["ghf"].cuda()).sum()
     + (ln2 * w["gln"].cuda()).sum()).backward()
    eJ = (gJ.grad.cpu() - nJ.grad).abs().amax(dim=(0, 2)); eh = (gh.grad.cpu() - nh.grad).abs().amax(dim=(0, 2))
    print(which, "fwd msg err %.1e" % float((Jf2.detach().cpu() - Jf.detach()).abs().max()), "| err per t: J", ["%.1e" % x for x in eJ.tolist()], "h", ["%.1e" % x for x in eh.tolist()])
    if which == ["gJp", "ghp", "gJf", "ghf", "gln"]:
        print("  vs reference golden: J %.1e h %.1e (cpu autograd vs golden)" % (float((nJ.grad - t64(d["filt_gJ"])).abs().max()), float((nh.grad - t64(d["filt_gh"])).abs().max())))
