"""CPU: the restatements in oracle/ and the device-agnostic torch mirrors against fixtures produced by RUNNING THE
REFERENCE'S OWN PYTHON (tests/golden/make_golden_py.py through oracle/build_ref_py.py): svae/models/gmm.py:12-16,62-128,
svae/models/iid_gausisan.py:42-46, svae/models/slds_svae.py:80-175,203-310.  These pin oracle/gmm_numpy.py and
oracle/slds_numpy.py (round 1: "parity unpinned") to outputs of the reference itself."""
import numpy as np
import pytest
import torch
from util import golden_cases, load, rel_err, rel_err_elem, slds_global_from_fixture

GMM = golden_cases("gmm_")
SLDS = golden_cases("slds_")
t64 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)


@pytest.mark.parametrize("path", GMM, ids=[p.split("/")[-1][:-4] for p in GMM])
def test_gmm_oracle_is_pinned_to_the_reference(path):
    from oracle import gmm_numpy as G
    d = load(path)
    local, prior, nat, kl, iters = G.local_meanfield((d["glob_dir"], d["glob_niw"]), (d["node_J"], d["node_h"]), d["label_init"])
    assert iters == int(d["n_sweeps"])
    assert rel_err_elem(local[0], d["label_stats"]) < 1e-9 and rel_err_elem(local[1], d["gaussian_stats"]) < 1e-9
    assert rel_err_elem(prior[0], d["dir_stats"]) < 1e-9 and rel_err_elem(prior[1], d["niw_stats"]) < 1e-9
    assert rel_err_elem(nat[0], d["label_natparam"]) < 1e-9 and rel_err_elem(nat[1], d["gaussian_natparam"]) < 1e-9
    assert abs(kl - float(d["kl"])) < 1e-10 * abs(float(d["kl"]))


@pytest.mark.parametrize("path", GMM, ids=[p.split("/")[-1][:-4] for p in GMM])
def test_gmm_torch_mirror_matches_the_reference(path):
    """svae_b200/models/gmm.py on the CPU (the CUDA fixed-point kernel has its own GPU test)."""
    from svae_b200.models import gmm
    d = load(path)
    glob, prior = (t64(d["glob_dir"]), t64(d["glob_niw"])), (t64(d["prior_dir"]), t64(d["prior_niw"]))
    pots = (t64(d["node_J"]), t64(d["node_h"]))
    local, pstats, nat, kl = gmm.local_meanfield(glob, pots, label_init=t64(d["label_init"]))
    c = lambda x: x.detach().numpy()
    assert rel_err_elem(c(local[0]), d["label_stats"]) < 1e-8 and rel_err_elem(c(local[1]), d["gaussian_stats"]) < 1e-8
    assert rel_err_elem(c(pstats[0]), d["dir_stats"]) < 1e-8 and rel_err_elem(c(pstats[1]), d["niw_stats"]) < 1e-8
    assert rel_err_elem(c(nat[1]), d["gaussian_natparam"]) < 1e-8
    assert abs(float(kl) - float(d["kl"])) < 1e-9 * abs(float(d["kl"]))
    S = d["eps"].shape[-1]
    samples, stats, global_kl, local_kl = gmm.run_inference(prior, glob, pots, S, label_init=t64(d["label_init"]),
                                                            eps=t64(d["eps"]))
    assert rel_err_elem(c(samples), d["samples"]) < 1e-8
    assert rel_err_elem(c(stats[0]), d["run_label_stats"]) < 1e-8 and rel_err_elem(c(stats[1]), d["run_niw_stats"]) < 1e-8
    assert abs(float(local_kl) - float(d["local_kl"])) < 1e-9 * abs(float(d["local_kl"]))
    # prior_kl: the intended inner product over the whole flattened parameter (the reference's checked-in util.flat
    # is shadowed and multiplies first elements only -- see make_golden_py.py; both values are in the fixture)
    assert abs(float(global_kl) - float(d["global_kl"])) < 1e-9 * abs(float(d["global_kl"]))
    assert abs(float(d["global_kl_as_checked_in"]) - float(d["global_kl"])) > 1.0


def test_iid_gaussian_matches_the_reference():
    """svae/models/iid_gausisan.py:42-46 (a13)."""
    from svae_b200.models import iid_gaussian as iid
    (path,) = golden_cases("iid_")
    d = load(path)
    S = d["eps"].shape[1]
    samples, stats, gvlb, lvlb = iid.run_inference(None, None, (t64(d["neghalfJ"]), t64(d["h"]), None), S, eps=t64(d["eps"]))
    assert stats == () and gvlb == 0.
    assert rel_err_elem(samples.numpy(), d["samples"]) < 1e-12
    assert abs(float(lvlb) - float(d["local_vlb"])) < 1e-12 * abs(float(d["local_vlb"]))
    # differentiable in the potentials (the recognition network trains through it)
    J, h = t64(d["neghalfJ"]).requires_grad_(True), t64(d["h"]).requires_grad_(True)
    s2, _, _, l2 = iid.run_inference(None, None, (J, h, None), S, eps=t64(d["eps"]))
    (s2.sum() + l2).backward()
    assert torch.isfinite(J.grad).all() and torch.isfinite(h.grad).all()
    # unsampled call: (T, n)
    assert iid.sample((t64(d["neghalfJ"]), t64(d["h"]), None)).shape == d["h"].shape


@pytest.mark.parametrize("path", SLDS, ids=[p.split("/")[-1][:-4] for p in SLDS])
def test_slds_oracle_is_pinned_to_the_reference(path):
    """oracle/slds_numpy.py, one sequence per call like the reference: same number of sweeps, same optimum."""
    from oracle import slds_numpy as SO
    d = load(path)
    glob = slds_global_from_fixture(d)
    B = d["node_J"].shape[0]
    for b in range(B if d["node_J"].shape[1] < 100 else 1):       # the T = 200 fixture: one sequence keeps the CPU suite short
        node = (d["node_J"][b:b + 1], d["node_h"][b:b + 1])
        hmm, (ist, pst, en), _, (hv, lv), it = SO.optimize_local_meanfield(glob, node, d["init_eps"][b:b + 1])
        assert it == int(d["n_sweeps"][b])
        assert rel_err_elem(hmm[2][0], d["Ez"][b]) < 1e-8 and rel_err_elem(hmm[1][0], d["Etrans"][b]) < 1e-8
        assert rel_err_elem(hmm[0][0], d["E0"][b]) < 1e-8
        assert rel_err_elem(ist[0][0], d["init_xx"][b]) < 1e-8 and rel_err_elem(pst[1][0], d["pair_xxn"][b], 1e-7) < 1e-7
        assert abs(hv[0] - d["hmm_vlb"][b]) < 1e-9 * abs(d["hmm_vlb"][b]) and abs(lv[0] - d["lds_vlb"][b]) < 1e-9 * abs(d["lds_vlb"][b])
        assert rel_err_elem(SO.arhmm_node(glob[1], tuple(x[b:b + 1] for x in (d["init_xx"], d["init_x"])),
                                          tuple(x[b:b + 1] for x in (d["pair_xx"], d["pair_xxn"], d["pair_xnxn"])))[0],
                            d["node_pot_hmm"][b]) < 1e-9


def test_slds_batched_oracle_equals_per_sequence_reference_runs():
    """B sequences in one call of the restatement == B reference calls (sweep counts 7 / 9 / 5 in the fixture)."""
    from oracle import slds_numpy as SO
    d = load([p for p in SLDS if "K3_n3_T12" in p][0])
    glob = slds_global_from_fixture(d)
    hmm, (ist, pst, en), _, (hv, lv), iters = SO.optimize_local_meanfield(glob, (d["node_J"], d["node_h"]), d["init_eps"])
    assert list(iters) == [int(x) for x in d["n_sweeps"]] and len(set(iters)) > 1
    assert rel_err_elem(hmm[2], d["Ez"]) < 1e-8 and rel_err_elem(pst[1], d["pair_xxn"], 1e-7) < 1e-7
    assert np.abs(hv - d["hmm_vlb"]).max() < 1e-9 and np.abs(lv - d["lds_vlb"]).max() < 1e-9


def test_live_reference_python_reproduces_a_fixture():
    """With /root/reference present (build container): stage the reference's Python again and re-run one GMM case."""
    from oracle import build_ref_py as R
    if not R.available():
        pytest.skip("/root/reference is absent (GPU box): fixtures only")
    d = load([p for p in GMM if "K4_N5_T33" in p][0])
    with R.staged() as ref:
        ref.gmm.initialize_meanfield = lambda lg, pots: d["label_init"]
        (label_stats, gaussian_stats), _, _, kl = ref.gmm.local_meanfield((d["glob_dir"], d["glob_niw"]), (d["node_J"], d["node_h"]))
    assert rel_err_elem(label_stats, d["label_stats"]) < 1e-12 and abs(kl - float(d["kl"])) < 1e-12 * abs(kl)
