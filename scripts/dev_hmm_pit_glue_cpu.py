"""Dev: run svae_b200/hmm/hmm_pit.py's glue (E-step and Viterbi) on CPU with the kernel entry points replaced by the
host builds of their phase code / the oracle, to check indexing before a GPU sees it."""
import sys, os, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
from oracle import hmm_numpy as HO
import hmm_pit_numpy as HP
from svae_b200 import _lib
from svae_b200.hmm import hmm_pit, hmm_inference as H

HP.stage12_host(*HO.rand_hmm_natparam(2, np.random.default_rng(0)), np.zeros((1, 3, 2)), 1)     # builds the host library
host = HP._host
T64 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)


class FakeLib:
    def svae_lds_compose_num_chunks(self, T, L):
        return (T - 1 + L - 1) // L

    def svae_hmm_compose_chunks(self, B, T, K, Lc, lA, node, P, ls, st):
        host.hmm_compose_chunks_host(B, T, K, Lc, lA, node, P, ls); return 0

    def svae_hmm_chain_messages(self, B, C, K, lp, P, ls, last, la, lb, lz, st):
        host.hmm_chain_messages_host(B, C, K, lp, P, ls, last, la, lb, lz); return 0

    def svae_hmm_compose_chunks_maxplus(self, B, T, K, Lc, lA, node, V, st):
        host.hmm_compose_chunks_maxplus_host(B, T, K, Lc, lA, node, V); return 0

    def svae_hmm_chain_viterbi(self, B, C, K, lp, V, last, delta, zb, st):
        host.hmm_chain_viterbi_host(B, C, K, lp, V, last, delta, zb); return 0


_lib.lib = lambda: FakeLib()
_lib.ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
_lib.stream_ptr = lambda dev: None
_lib.f64 = lambda x, dev=None: (x if torch.is_tensor(x) else torch.as_tensor(x, dtype=torch.float64)).to(torch.float64).contiguous()
H._canon.__globals__["torch"].cuda.current_device = lambda: 0
_orig_canon = H._canon


def canon(natparam):
    lp, lA, node = (_lib.f64(x) for x in natparam)
    batched = node.dim() == 3
    if not batched:
        node = node[None]
    return lp, lA, node.contiguous(), batched, torch.device("cpu")


H._canon = canon


def fake_estep_forward(lp, lA, node, logz_only=False):
    logZ, (E0, Et, Ez) = HO.hmm_estep((lp.numpy(), lA.numpy(), node.numpy()))
    return T64(logZ), T64(E0), T64(Et), T64(Ez)


def fake_viterbi(natparam):
    lp, lA, node = (x.numpy() for x in natparam)
    return torch.as_tensor(HO.hmm_viterbi((lp, lA, node)), dtype=torch.int32)


H.hmm_estep_forward = fake_estep_forward
H.hmm_viterbi = fake_viterbi
err = lambda a, b: float(np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-300))
for K, T, B, Lc in [(3, 10, 2, 3), (8, 41, 3, 7), (32, 60, 2, None), (17, 53, 2, 16), (5, 9, 1, 1)]:
    rng = np.random.default_rng(K + T)
    lp, lA = HO.rand_hmm_natparam(K, rng)
    node = 2.0 * rng.standard_normal((B, T, K))
    w_logZ, (w0, wt, wz) = HO.hmm_estep((lp, lA, node))
    logZ, (E0, Et, Ez) = hmm_pit.hmm_estep_parallel((T64(lp), T64(lA), T64(node)), chunk_len=Lc)
    path = hmm_pit.hmm_viterbi_parallel((T64(lp), T64(lA), T64(node)), chunk_len=Lc)
    print((K, T, B, Lc), "logZ %.1e Ez %.1e Etrans %.1e" % (err(logZ, w_logZ), err(Ez, wz), err(Et, wt)),
          "viterbi equal:", bool(np.array_equal(path.numpy(), HO.hmm_viterbi((lp, lA, node)))))
