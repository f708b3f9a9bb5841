"""Checkpoint stream: the format experiments/gmm_svae_synth_plot_pdfs.py:201-208 reads (records until EOFError)."""
import pickle
import numpy as np
import torch
from svae_b200 import io


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, dtype=torch.float64, generator=g)
    pgm = (r(5), (r(5, 4, 4), r(5, 4)))
    mlp = [(r(2, 8), r(8)), (r(8, 4), r(4))]
    return (pgm, mlp, list(mlp))


def test_stream_roundtrip_and_reference_reader(tmp_path):
    fn = str(tmp_path / "gmm_svae_synth_params.pkl")
    for i in range(3):
        io.append_params(fn, _params(i))
    # the reference's reader, verbatim in spirit: load until EOFError, keep the last
    with open(fn, "rb") as f:
        try:
            for _ in range(20000):
                last = pickle.load(f)
        except EOFError:
            pass
    want = io.to_numpy_tree(_params(2))
    assert isinstance(last, tuple) and isinstance(last[1], list) and isinstance(last[0][1][0], np.ndarray)
    assert np.array_equal(last[0][1][0], want[0][1][0]) and np.array_equal(last[2][1][0], want[2][1][0])
    back = io.load_last_params(fn)
    assert torch.equal(back[0][0], _params(2)[0][0]) and back[1][0][0].dtype == torch.float64
    assert len(list(io.iter_params(fn))) == 3


def test_truncated_trailing_record_is_ignored(tmp_path):
    fn = str(tmp_path / "p.pkl")
    io.append_params(fn, _params(0))
    io.append_params(fn, _params(1))
    size = __import__("os").path.getsize(fn)
    with open(fn, "r+b") as f:
        f.truncate(size - 7)              # a crash in the middle of the second dump
    back = io.load_last_params(fn)
    assert torch.equal(back[0][0], _params(0)[0][0])
