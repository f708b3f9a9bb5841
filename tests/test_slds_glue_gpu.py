"""GPU: the two SLDS coupling contractions behind the C-ABI (svae_slds_node_potentials, svae_slds_mix_params;
svae/models/slds_svae.py:92-103,131-147) against torch einsum in fp64 -- the fp64 tensor-core path (K <= 8) and the
general path (K > 8), ragged row counts, odd n, T = 1, frozen sequences."""
import ctypes
import pytest
import torch

pytestmark = pytest.mark.gpu


def _problem(B, T, n, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.randn(*s, dtype=torch.float64, device="cuda", generator=g)
    Tm = max(T - 1, 1)
    d = dict(Exx=r(B, T, n, n), Exxn=r(B, Tm, n, n), Ex0=r(B, n), iJ=r(K, n, n), ih=r(K, n), ic=r(K),
             J11=r(K, n, n), J12=r(K, n, n), J22=r(K, n, n), lZ=r(K), Ez=torch.softmax(r(B, T, K), -1))
    d["active"] = (torch.rand(B, device="cuda", generator=g) > 0.3).to(torch.int32)
    return d


@pytest.mark.parametrize("B,T,n,K", [(5, 13, 10, 8), (3, 7, 5, 3), (4, 1, 4, 8), (2, 9, 3, 12), (37, 50, 10, 8), (1, 3, 16, 8)])
@pytest.mark.parametrize("masked", [False, True])
def test_node_potentials_and_mix_params(B, T, n, K, masked):
    from svae_b200 import _lib
    d = _problem(B, T, n, K, 100 * n + K)
    L, P = _lib.lib(), _lib.ptr
    st = _lib.stream_ptr(torch.device("cuda"))
    act = d["active"] if masked else None
    ap = ctypes.c_void_p(act.data_ptr()) if masked else None
    keep = (act.bool() if masked else torch.ones(B, dtype=torch.bool, device="cuda"))
    sentinel = -7.25
    node = torch.full((B, T, K), sentinel, dtype=torch.float64, device="cuda")
    _lib.check(L.svae_slds_node_potentials(B, T, n, K, P(d["Exx"]), P(d["Exxn"]), P(d["Ex0"]), P(d["iJ"]), P(d["ih"]), P(d["ic"]),
                                           P(d["J11"]), P(d["J12"]), P(d["J22"]), P(d["lZ"]), ap, P(node), st), "node")
    want = torch.empty_like(node)
    want[:, 0] = torch.einsum("bij,kij->bk", d["Exx"][:, 0], d["iJ"]) + d["Ex0"] @ d["ih"].T + d["ic"]
    if T > 1:
        want[:, 1:] = (torch.einsum("btij,kij->btk", d["Exx"][:, :-1], d["J11"]) + torch.einsum("btij,kij->btk", d["Exxn"], d["J12"])
                       + torch.einsum("btij,kij->btk", d["Exx"][:, 1:], d["J22"]) + d["lZ"])
    assert torch.allclose(node[keep], want[keep], rtol=1e-12, atol=1e-12)
    assert (node[~keep] == sentinel).all()            # frozen sequences keep what they had

    Tm = max(T - 1, 1)
    new = lambda *s: torch.full(s, sentinel, dtype=torch.float64, device="cuda")
    oJ, oh, oc = new(B, n, n), new(B, n), new(B)
    o11, o12, o22, olZ = new(B, Tm, n, n), new(B, Tm, n, n), new(B, Tm, n, n), new(B, Tm)
    _lib.check(L.svae_slds_mix_params(B, T, n, K, P(d["Ez"]), P(d["iJ"]), P(d["ih"]), P(d["ic"]), P(d["J11"]), P(d["J12"]),
                                      P(d["J22"]), P(d["lZ"]), ap, P(oJ), P(oh), P(oc), P(o11), P(o12), P(o22), P(olZ), st), "mix")
    Ez = d["Ez"]
    close = lambda a, b: torch.allclose(a[keep], b[keep], rtol=1e-12, atol=1e-13)
    assert close(oJ, torch.einsum("bk,kij->bij", Ez[:, 0], d["iJ"]))
    assert close(oh, Ez[:, 0] @ d["ih"]) and close(oc, Ez[:, 0] @ d["ic"])
    if T > 1:
        for got, par in ((o11, "J11"), (o12, "J12"), (o22, "J22")):
            assert close(got, torch.einsum("btk,kij->btij", Ez[:, 1:], d[par]))
        assert close(olZ, Ez[:, 1:] @ d["lZ"])
    for x in (oJ, oh, oc) + ((o11, o12, o22, olZ) if T > 1 else ()):
        assert (x[~keep] == sentinel).all()
