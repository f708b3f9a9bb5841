"""GPU: the CUDA paths of a11 (GMM), a12 (SLDS) and a13 (iid) against fixtures produced by RUNNING THE REFERENCE'S
OWN PYTHON (tests/golden/make_golden_py.py).  SLDS at the BASELINE shape K = 8, d = 10, T = 200 included, with the
gradient w.r.t. the node potentials checked against torch autograd through the CPU oracle."""
import numpy as np
import pytest
import torch
from util import golden_cases, load, rel_err, rel_err_elem, slds_global_from_fixture

pytestmark = pytest.mark.gpu
GMM = golden_cases("gmm_")
SLDS = golden_cases("slds_")
g64 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")
c = lambda x: x.detach().cpu().numpy()


@pytest.mark.parametrize("path", GMM, ids=[p.split("/")[-1][:-4] for p in GMM])
def test_gmm_cuda_matches_the_reference(path):
    """svae_gmm_meanfield (the whole fixed point of gmm.py:90-110 in one launch) + the evaluation at the optimum."""
    from svae_b200.models import gmm
    from svae_b200.distributions import dirichlet, niw
    d = load(path)
    glob, prior = (g64(d["glob_dir"]), g64(d["glob_niw"])), (g64(d["prior_dir"]), g64(d["prior_niw"]))
    pots = (g64(d["node_J"]), g64(d["node_h"]))
    labels, iters, kl = gmm.meanfield_fixed_point_cuda(dirichlet.expectedstats(glob[0]), niw.expectedstats(glob[1]),
                                                       pots[0], pots[1], g64(d["label_init"]))
    assert int(iters) == int(d["n_sweeps"])
    local, pstats, nat, kl = gmm.local_meanfield(glob, pots, label_init=g64(d["label_init"]))
    assert rel_err_elem(c(local[0]), d["label_stats"]) < 1e-8 and rel_err_elem(c(local[1]), d["gaussian_stats"]) < 1e-8
    assert rel_err_elem(c(pstats[0]), d["dir_stats"]) < 1e-8 and rel_err_elem(c(pstats[1]), d["niw_stats"]) < 1e-8
    assert abs(float(kl) - float(d["kl"])) < 1e-9 * abs(float(d["kl"]))
    S = d["eps"].shape[-1]
    samples, stats, global_kl, local_kl = gmm.run_inference(prior, glob, pots, S, label_init=g64(d["label_init"]), eps=g64(d["eps"]))
    assert rel_err_elem(c(samples), d["samples"]) < 1e-8
    assert abs(float(local_kl) - float(d["local_kl"])) < 1e-9 * abs(float(d["local_kl"]))
    assert abs(float(global_kl) - float(d["global_kl"])) < 1e-9 * abs(float(d["global_kl"]))


def test_iid_gaussian_cuda_matches_the_reference():
    from svae_b200.models import iid_gaussian as iid
    (path,) = golden_cases("iid_")
    d = load(path)
    samples, _, _, lvlb = iid.run_inference(None, None, (g64(d["neghalfJ"]), g64(d["h"]), None), d["eps"].shape[1], eps=g64(d["eps"]))
    assert rel_err_elem(c(samples), d["samples"]) < 1e-12
    assert abs(float(lvlb) - float(d["local_vlb"])) < 1e-12 * abs(float(d["local_vlb"]))


def _cuda_global(d, which):
    (dirn, mdir), lds = slds_global_from_fixture(d, which)
    return ((g64(dirn), g64(mdir)), [(g64(a), tuple(g64(x) for x in b)) for a, b in lds])


def _check_slds(d, rows, out, tol=1e-7):
    samples, stats, global_vlb, local_vlb = out
    (E0, Etrans), lds_g = stats
    K = d["glob_dir"].shape[0]
    assert rel_err_elem(c(samples), d["samples"][rows], 1e-7) < tol
    assert abs(float(local_vlb) - d["local_vlb"][rows].sum()) < 1e-8 * abs(d["local_vlb"][rows].sum())
    assert abs(float(global_vlb) - float(d["global_vlb"])) < 1e-9 * abs(float(d["global_vlb"]))
    assert rel_err_elem(c(E0), d["g_E0"][rows].sum(0)) < tol and rel_err_elem(c(Etrans), d["g_Etrans"][rows].sum(0)) < tol
    from svae_b200.distributions import gaussian
    for k in range(K):
        xx, x, a, b = (c(v) for v in gaussian.unpack_dense(lds_g[k][0]))
        assert rel_err_elem(xx, d["g_init_xx"][rows][:, k].sum(0), 1e-6) < tol and rel_err_elem(x, d["g_init_x"][rows][:, k].sum(0), 1e-6) < tol
        assert abs(a - d["g_init_a"][rows][:, k].sum()) < tol and abs(b - d["g_init_b"][rows][:, k].sum()) < tol
        for j, key in enumerate(("g_pair_xx", "g_pair_xxn", "g_pair_xnxn", "g_pair_n")):
            assert rel_err_elem(c(lds_g[k][1][j]), d[key][rows][:, k].sum(0), 1e-6) < tol, (k, key)


@pytest.mark.parametrize("path", SLDS, ids=[p.split("/")[-1][:-4] for p in SLDS])
def test_slds_run_inference_matches_the_reference_one_sequence_per_call(path):
    """slds_svae.py:289-310 as the reference runs it: one sequence per call (B = 1)."""
    from svae_b200.models import slds_svae as M
    d = load(path)
    glob, prior = _cuda_global(d, "glob"), _cuda_global(d, "prior")
    B, S = d["node_J"].shape[0], d["eps"].shape[2]
    for b in range(B):
        rows = slice(b, b + 1)
        out = M.run_inference(prior, glob, (g64(d["node_J"][rows]), g64(d["node_h"][rows])), S, eps=g64(d["eps"][rows]),
                              init_eps=g64(d["init_eps"][rows]))
        _check_slds(d, rows, out)
        (hmm_stats, lds_stats), _, (hv, lv) = M.optimize_local_meanfield(
            glob, (g64(d["node_J"][rows]), g64(d["node_h"][rows])), init_eps=g64(d["init_eps"][rows]))
        assert rel_err_elem(c(hmm_stats[2][0]), d["Ez"][b]) < 1e-7 and rel_err_elem(c(hmm_stats[1][0]), d["Etrans"][b]) < 1e-7
        assert abs(float(hv[0]) - d["hmm_vlb"][b]) < 1e-8 * abs(d["hmm_vlb"][b])
        assert abs(float(lv[0]) - d["lds_vlb"][b]) < 1e-8 * abs(d["lds_vlb"][b])


@pytest.mark.parametrize("path", SLDS, ids=[p.split("/")[-1][:-4] for p in SLDS])
def test_slds_run_inference_batched_equals_the_per_sequence_reference(path):
    """B sequences in one call: every sequence stops when ITS bound has converged (the reference's rule per sequence),
    so the batch reproduces B independent reference runs whatever their sweep counts (here 7/9/5 and 5/5/5/5)."""
    from svae_b200.models import slds_svae as M
    d = load(path)
    glob, prior = _cuda_global(d, "glob"), _cuda_global(d, "prior")
    S = d["eps"].shape[2]
    out = M.run_inference(prior, glob, (g64(d["node_J"]), g64(d["node_h"])), S, eps=g64(d["eps"]), init_eps=g64(d["init_eps"]))
    _check_slds(d, slice(None), out)


def test_slds_gradient_at_the_baseline_shape():
    """K = 8, d = 10, T = 200: d(local_vlb + <w, samples>) / d(node potentials) from the CUDA VJP kernels against torch
    autograd through the CPU oracle of the differentiable tail (slds_svae.py:297-310 at the mean-field optimum)."""
    from svae_b200.models import slds_svae as M
    from oracle import lds_torch as OT
    d = load([p for p in SLDS if "K8_n10_T200" in p][0])
    glob, prior = _cuda_global(d, "glob"), _cuda_global(d, "prior")
    rows = slice(0, 1)
    nJ, nh = g64(d["node_J"][rows]).requires_grad_(True), g64(d["node_h"][rows]).requires_grad_(True)
    w = torch.randn(d["samples"][rows].shape, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    samples, _, _, local_vlb = M.run_inference(prior, glob, (nJ, nh), 1, eps=g64(d["eps"][rows]), init_eps=g64(d["init_eps"][rows]))
    (local_vlb + (samples * w.cuda()).sum()).backward()
    # the CPU oracle of the tail: fixed local natural parameters (computed from unboxed potentials), torch autograd
    with torch.no_grad():
        (hmm_stats, _), (hmm_nat, lds_nat), _ = M.optimize_local_meanfield(glob, (nJ.detach(), nh.detach()), init_eps=g64(d["init_eps"][rows]))
        dense_init, dense_pair = M.get_all_lds_local_natparams(glob[1])
    cpu = lambda x: x.detach().cpu()
    init_p, pair = lds_nat
    init3 = (cpu(init_p[0]), cpu(init_p[1]), cpu(sum(init_p[2:])))
    pairc = tuple(cpu(p) for p in pair)
    Jn, hn = cpu(nJ).clone().requires_grad_(True), cpu(nh).clone().requires_grad_(True)
    smp, (E_init, E_pair, E_node), lognorm = OT.natural_lds_inference_general((init3, pairc), (Jn, hn), cpu(g64(d["eps"][rows])))
    (Jk, hk, ak, bk), (J11k, J12k, J22k, lZk) = tuple(cpu(x) for x in dense_init), tuple(cpu(x) for x in dense_pair)
    init_node = torch.einsum('bij,kij->bk', E_init[0], Jk) + E_init[1] @ hk.T + (ak + bk)[None]
    rest = torch.einsum('btij,kij->btk', E_pair[0], J11k) + torch.einsum('btij,kij->btk', E_pair[1], J12k) \
        + torch.einsum('btij,kij->btk', E_pair[2], J22k) + lZk[None, None]
    node = torch.cat((init_node[:, None], rest), 1)[0]
    la = cpu(hmm_nat[0]) + node[0]
    for t in range(1, node.shape[0]):
        la = torch.logsumexp(la[:, None] + cpu(hmm_nat[1]), 0) + node[t]
    want_vlb = torch.logsumexp(la, 0) + lognorm.sum() - (Jn * E_node[0]).sum() - (hn * E_node[1]).sum()
    assert abs(float(want_vlb) - float(local_vlb)) < 1e-8 * abs(float(want_vlb))
    (want_vlb + (smp * w).sum()).backward()
    assert rel_err_elem(c(nJ.grad), Jn.grad.numpy(), 1e-6) < 1e-6
    assert rel_err_elem(c(nh.grad), hn.grad.numpy(), 1e-6) < 1e-6
