"""Shared helpers for the parity tests."""
import glob, os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_cases(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    """max |a-b| / max |b|  (the '<= 1e-5 rel on posterior moments' metric of BASELINE.json)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def golden_natparams(d):
    init = (d["init_J"], d["init_h"], float(d["init_logZ"]))
    pair = (d["J11"], d["J12"], d["J22"], float(d["pair_logZ"]))
    node = (d["node_J"], d["node_h"], d["node_logZ"])
    return init, pair, node


def rel_err_elem(a, b, floor=1e-9):
    """max over elements of |a-b| / (|b| + floor * max|b|): a wrong SMALL element shows up (rel_err would hide it)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    return float((np.abs(a - b) / (np.abs(b) + floor * max(np.abs(b).max(), 1e-300))).max())


def slds_global_from_fixture(d, which="glob"):
    """((dir, mdir), [(niw packed, (A, B, C, d))] * K) from a tests/golden/slds_*.npz fixture (NumPy)."""
    K = d[which + "_dir"].shape[0]
    return ((d[which + "_dir"], d[which + "_mdir"]),
            [(d[which + "_niw"][k], tuple(d[which + "_mniw_" + c][k] for c in "ABCd")) for k in range(K)])
