"""Nested-tuple algebra and batching helpers -- torch mirror of the parts of the reference's svae/util.py that
the SVI loop and the E-step glue use (svae/util.py:16-24 activations, :59-78 small linear algebra, :105-125
batching, :130-174 arithmetic on nested tuples).  Leaves are torch tensors (or python scalars)."""
from __future__ import annotations
import math
import torch

istuple = lambda x: isinstance(x, (tuple, list))
isarray = lambda x: torch.is_tensor(x)

T = lambda X: X.transpose(-1, -2)
symmetrize = lambda X: (X + T(X)) / 2.
outer = lambda x, y: x[..., :, None] * y[..., None, :]
normalize = lambda x: x / x.sum(-1, keepdim=True)
softmax = lambda x: torch.softmax(x, dim=-1)
log1pexp = torch.nn.functional.softplus
sigmoid, relu, identity = torch.sigmoid, torch.relu, (lambda x: x)


def _leaf(x, like=None):
    if torch.is_tensor(x):
        return x
    return torch.as_tensor(x, dtype=torch.float64, device=None if like is None else like.device)


def _unop(op, combine):
    def f(a):
        return combine([f(x) for x in a]) if istuple(a) else op(a)
    return f


def _binop(op, combine):
    def f(a, b):
        if istuple(a):
            assert istuple(b) and len(a) == len(b), "nested structures differ"
            return combine([f(x, y) for x, y in zip(a, b)])
        return op(a, b)
    return f


def _scalar_op(op, combine):
    def f(a, b):
        return combine([f(a, y) for y in b]) if istuple(b) else op(a, b)
    return f


add = _binop(lambda a, b: a + b, tuple)
sub = _binop(lambda a, b: a - b, tuple)
mul = _binop(lambda a, b: a * b, tuple)
div = _binop(lambda a, b: a / b, tuple)
contract = _binop(lambda a, b: (_leaf(a) * _leaf(b)).sum(), lambda xs: sum(xs))     # util.py:149
scale = _scalar_op(lambda a, b: a * b, tuple)
add_scalar = _scalar_op(lambda a, b: a + b, tuple)
shape = _unop(lambda a: tuple(a.shape) if torch.is_tensor(a) else (), tuple)
unbox = _unop(lambda a: a.detach() if torch.is_tensor(a) else a, tuple)              # getval
zeros_like = _unop(lambda a: torch.zeros_like(a) if torch.is_tensor(a) else 0., tuple)
square = _unop(lambda a: a ** 2, tuple)
sqrt = _unop(lambda a: a ** 0.5, tuple)
allclose = _binop(lambda a, b: bool(torch.allclose(_leaf(a), _leaf(b, _leaf(a)))), all)


def flatten(struct):
    """autograd.util.flatten: (flat 1-D tensor, unflatten)."""
    leaves, spec = [], []

    def walk(a):
        if istuple(a):
            return tuple(walk(x) for x in a)
        t = _leaf(a)
        leaves.append(t.reshape(-1))
        return (len(leaves) - 1, tuple(t.shape))
    tree = walk(struct)
    flat_vec = torch.cat(leaves) if leaves else torch.zeros(0, dtype=torch.float64)
    sizes = [l.numel() for l in leaves]

    def unflatten(vec):
        parts = torch.split(vec, sizes)

        def build(node):
            if isinstance(node, tuple) and len(node) == 2 and isinstance(node[0], int) and isinstance(node[1], tuple) \
                    and all(isinstance(d, int) for d in node[1]):
                return parts[node[0]].reshape(node[1])
            return tuple(build(x) for x in node)
        return build(tree)
    return flat_vec, unflatten


flat = lambda struct: flatten(struct)[0]
norm = lambda x: contract(x, x) ** 0.5


def rand_psd(n, generator=None, device=None):
    A = torch.randn(n, n, dtype=torch.float64, generator=generator, device=device)
    return A @ A.T


# ---- batching (util.py:105-125) -------------------------------------------------------------------

def get_num_datapoints(x):
    return x.shape[0] if torch.is_tensor(x) else sum(get_num_datapoints(y) for y in x)


def split_into_batches(data, seq_len, generator=None):
    """(N,p) data -> list of (seq_len,p) contiguous chunks, randomly permuted (util.py:115-125).
    A 3-D tensor (num_sequences, T, p) is instead split into minibatches of `seq_len` SEQUENCES -- the batch
    axis this build adds."""
    if torch.is_tensor(data) and data.dim() == 3:
        k = data.shape[0] // seq_len
        chunks = list(torch.split(data[:k * seq_len], seq_len))
    else:
        arrs = [data] if torch.is_tensor(data) else list(data)
        chunks = []
        for arr in arrs:
            k = arr.shape[0] // seq_len
            chunks += list(torch.split(arr[:k * seq_len], seq_len))
    perm = torch.randperm(len(chunks), generator=generator).tolist()
    chunks = [chunks[i] for i in perm]
    return chunks, len(chunks)


# ---- small dense linear algebra without a device->host round trip ----------------------------------------------
# torch.linalg.inv / solve / cholesky read their LAPACK-style info word on the host to raise on singular input: on a
# CUDA stream that is a synchronisation per call (the ~40 tiny exp-family maps of an SVI step would each stall the
# launch queue).  The *_ex variants leave the check to the caller; on CUDA a singular / non-PD input then shows up
# as inf / nan in the result (and in the kernels' `status` word), on the CPU the checked versions are kept.

def inv(X):
    return torch.linalg.inv_ex(X).inverse if X.is_cuda else torch.linalg.inv(X)


def solve(A, B):
    return torch.linalg.solve_ex(A, B).result if A.is_cuda else torch.linalg.solve(A, B)


def cholesky(X):
    return torch.linalg.cholesky_ex(X).L if X.is_cuda else torch.linalg.cholesky(X)
