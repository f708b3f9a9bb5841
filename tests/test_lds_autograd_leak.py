"""The differentiable LDS E-step must not keep its outputs alive after backward (ctx -> outputs -> grad_fn -> ctx
would be a cycle the collector cannot break: ~260 MB per SVI step at the metric shape).  Kernels are mocked: this is
host logic and runs without a GPU."""
import gc, weakref
import torch
from svae_b200.lds import lds_inference as K


def _fake_forward(init_params, pair_params, node, S, eps, per_t_pair_stats=False, save_for_bwd=False, **kw):
    nJ, nh = node[0], node[1]
    B, T, n = nh.shape
    z = lambda *s: torch.zeros(s, dtype=torch.float64)
    o = dict(B=B, T=T, n=n, S=S, batched=True, eps=eps, dense=False, lognorm=z(B), E_node_xx=z(B, T, n), E_node_x=z(B, T, n),
             samples=z(B, T, S, n) if S else None, E_init_xx=z(B, n, n), E_init_x=z(B, n), E_pair_xx=z(B, n, n),
             E_pair_xxn=z(B, n, n), E_pair_xnxn=z(B, n, n), status=torch.zeros(1, dtype=torch.int32))
    if save_for_bwd:
        o["workspace"], o["workspace_bytes"] = torch.zeros(64, dtype=torch.uint8), 64
    return o


def _fake_backward(o, g):
    assert o["workspace"] is not None and o["E_node_x"] is not None
    B, T, n = o["B"], o["T"], o["n"]
    return torch.ones(B, T, n, dtype=torch.float64), torch.ones(B, T, n, dtype=torch.float64)


def test_outputs_are_freed_after_backward(monkeypatch):
    monkeypatch.setattr(K, "lds_estep_forward", _fake_forward)
    monkeypatch.setattr(K, "lds_estep_backward", _fake_backward)
    B, T, n, S = 3, 5, 2, 1
    nJ = torch.zeros(B, T, n, dtype=torch.float64, requires_grad=True)
    nh = torch.zeros(B, T, n, dtype=torch.float64, requires_grad=True)
    eps = torch.zeros(B, T, S, n, dtype=torch.float64)
    outs = K._LDSEStepFn.apply(nJ, nh, None, eps, None, None, S, False)
    refs = [weakref.ref(t) for t in outs] + [weakref.ref(eps)]
    (outs[0].sum() + outs[1].sum() + outs[3].sum()).backward()
    assert float(nJ.grad.sum()) == B * T * n
    del outs, eps
    gc.collect()
    assert all(r() is None for r in refs), [r() is not None for r in refs]


def test_outputs_are_freed_without_backward(monkeypatch):
    monkeypatch.setattr(K, "lds_estep_forward", _fake_forward)
    monkeypatch.setattr(K, "lds_estep_backward", _fake_backward)
    nJ = torch.zeros(2, 4, 2, dtype=torch.float64, requires_grad=True)
    nh = torch.zeros(2, 4, 2, dtype=torch.float64, requires_grad=True)
    outs = K._LDSEStepFn.apply(nJ, nh, None, None, None, None, 0, False)
    refs = [weakref.ref(t) for t in outs]
    del outs
    gc.collect()
    assert all(r() is None for r in refs)
