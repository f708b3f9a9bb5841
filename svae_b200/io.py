"""Parameter checkpoints as the pickle STREAM the reference's plotting scripts read: one `pickle.dump(params)` per
saved iteration appended to a single file, the last record being the current state
(experiments/gmm_svae_synth_plot_pdfs.py:201-208 reads records until EOFError and keeps the last).

`params` is the nested tuple/list the SVI loop carries, (pgm_params, loglike_params, recogn_params) (svae/svae.py:26);
tensors are written as NumPy arrays (protocol 2, loadable from the reference's Python 2) and come back as fp64
tensors on the requested device.
"""
from __future__ import annotations
import pickle
import numpy as np
import torch


def to_numpy_tree(x):
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    if isinstance(x, tuple):
        return tuple(to_numpy_tree(v) for v in x)
    if isinstance(x, list):
        return [to_numpy_tree(v) for v in x]
    if isinstance(x, dict):
        return {k: to_numpy_tree(v) for k, v in x.items()}
    return x


def to_tensor_tree(x, device=None, dtype=torch.float64):
    if isinstance(x, np.ndarray):
        return torch.as_tensor(x, dtype=dtype if x.dtype.kind == "f" else None, device=device)
    if isinstance(x, tuple):
        return tuple(to_tensor_tree(v, device, dtype) for v in x)
    if isinstance(x, list):
        return [to_tensor_tree(v, device, dtype) for v in x]
    if isinstance(x, dict):
        return {k: to_tensor_tree(v, device, dtype) for k, v in x.items()}
    return x


def append_params(filename, params):
    """Append one record (call it from make_gradfun's `callback`, svae/svae.py:36)."""
    with open(filename, "ab") as f:
        pickle.dump(to_numpy_tree(params), f, protocol=2)


def iter_params(filename, device=None):
    with open(filename, "rb") as f:
        while True:
            try:
                rec = pickle.load(f)
            except EOFError:
                return
            yield to_tensor_tree(rec, device)


def load_last_params(filename, device=None):
    """The resume point: the last complete record of the stream (a truncated trailing record is ignored)."""
    last = None
    try:
        for rec in iter_params(filename, device):
            last = rec
    except (pickle.UnpicklingError, AttributeError, ValueError, IndexError):
        pass
    if last is None:
        raise EOFError("no complete record in %s" % filename)
    return last
