"""ctypes binding of libsvae_b200.so (the C-ABI declared in include/svae_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every tensor crosses the boundary as a
raw device pointer.  There is NO CPU fallback: if the CUDA library is missing or a tensor is not a
CUDA fp64 tensor the call raises.
"""
from __future__ import annotations
import ctypes, os
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVAE_B200_LIB") or os.path.join(_HERE, "libsvae_b200.so")

SVAE_LDS_MAX_N = 16
SVAE_LDS_PER_T_PAIR_STATS = 1
SVAE_LDS_SAVE_FOR_BWD = 2
SVAE_LDS_INIT_PER_SEQ = 4
SVAE_LDS_BWD_MESSAGES_OUT = 32
SVAE_LDS_BWD_MESSAGES_IN = 64
SVAE_HMM_MAX_K = 32          # register-resident kernels (lane = state)
SVAE_HMM_WIDE_MAX_K = 1024   # svae_hmm_estep_wide / svae_hmm_viterbi (thread = state)
SVAE_HMM_TC_MIN_K = 17       # svae_hmm_estep_wide: 17..32 states on the fp64 tensor-core instruction
SVAE_GMM_MAX_N = 8
SVAE_HMM_LOGZ_ONLY = 1

_ERR = {-1: "SVAE_ERR_BAD_ARG", -2: "SVAE_ERR_UNSUPPORTED", -3: "SVAE_ERR_WORKSPACE", -4: "SVAE_ERR_CUDA"}

_lib = None

_P = ctypes.c_void_p
_I = ctypes.c_int
_LL = ctypes.c_longlong
_SZ = ctypes.c_size_t

_SIGNATURES = {
    "svae_b200_version": (ctypes.c_char_p, []),
    "svae_b200_launch_count": (ctypes.c_uint64, []),
    "svae_lds_estep_workspace_bytes": (_SZ, [_I] * 5),
    "svae_lds_estep_fwd": (_I, [_I] * 5 + [_P] * 7 + [_LL, _LL] + [_P] * 15 + [_P, _P, _SZ, _P]),
    "svae_lds_estep_fwd_ex": (_I, [_I] * 5 + [_P] * 7 + [_LL, _LL] + [_P] * 21 + [_P, _P, _SZ, _P]),
    "svae_lds_estep_fwd_masked": (_I, [_I] * 5 + [_P] * 7 + [_LL, _LL] + [_P] * 22 + [_P, _P, _SZ, _P]),
    "svae_hmm_estep_masked": (_I, [_I] * 4 + [_P] * 8 + [_P]),
    "svae_hmm_wide_workspace_bytes": (_SZ, [_I] * 3),
    "svae_hmm_estep_wide": (_I, [_I] * 4 + [_P] * 8 + [_P, _SZ, _P]),
    "svae_slds_mix_params": (_I, [_I] * 4 + [_P] * 16 + [_P]),
    "svae_slds_node_potentials": (_I, [_I] * 4 + [_P] * 12 + [_P]),
    "svae_gaussian_info_head_fwd": (_I, [_LL, _I] + [_P] * 5 + [_P]),
    "svae_gaussian_info_head_bwd": (_I, [_LL, _I] + [_P] * 4 + [_P]),
    "svae_lds_compose_num_chunks": (_I, [_I, _I]),
    "svae_hmm_compose_chunks": (_I, [_I] * 4 + [_P] * 4 + [_P]),
    "svae_hmm_chain_messages": (_I, [_I] * 3 + [_P] * 7 + [_P]),
    "svae_hmm_compose_chunks_maxplus": (_I, [_I] * 4 + [_P] * 3 + [_P]),
    "svae_hmm_chain_viterbi": (_I, [_I] * 3 + [_P] * 5 + [_P]),
    "svae_lds_compose_chunks": (_I, [_I] * 4 + [_P] * 4 + [_LL, _LL] + [_P] * 9 + [_P, _P]),
    "svae_lds_estep_bwd_ex": (_I, [_I] * 5 + [_P] * 17 + [_P, _SZ, _P]),
    "svae_lds_estep_bwd_ports": (_I, [_I] * 5 + [_P] * 23 + [_P, _SZ, _P]),
    "svae_lds_estep_bwd": (_I, [_I] * 5 + [_P] * 16 + [_P, _SZ, _P]),
    "svae_lds_reduce_stats_workspace_bytes": (_SZ, [_I]),
    "svae_lds_reduce_stats": (_I, [_I] * 2 + [_P] * 7 + [_P, _SZ, _P]),
    "svae_hmm_estep": (_I, [_I] * 4 + [_P] * 7 + [_P]),
    "svae_gmm_meanfield": (_I, [_I] * 4 + [_P] * 5 + [ctypes.c_double, _I] + [_P] * 3 + [_P]),
    "svae_hmm_viterbi_workspace_bytes": (_SZ, [_I] * 3),
    "svae_hmm_viterbi": (_I, [_I] * 3 + [_P] * 4 + [_P, _SZ, _P]),
}


def lib() -> ctypes.CDLL:
    """Load (once) the CUDA library; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "svae_b200: %s is missing -- build it with `python -m svae_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(code: int, what: str):
    if code != 0:
        raise RuntimeError("%s failed: %s (%d)" % (what, _ERR.get(code, "?"), code))


def ptr(t: torch.Tensor | None):
    """Raw device pointer of a contiguous CUDA fp64 tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("svae_b200: expected a CUDA tensor, got device=%s (no CPU path)" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("svae_b200: tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def f64(x, device=None) -> torch.Tensor:
    """Contiguous CUDA fp64 view/copy of x."""
    if not torch.is_tensor(x):
        x = torch.as_tensor(x, dtype=torch.float64, device=device)
    if x.dtype != torch.float64:
        x = x.to(torch.float64)
    if device is not None and x.device != torch.device(device):
        x = x.to(device)
    return x.contiguous()


def stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count() -> int:
    return int(lib().svae_b200_launch_count())
