"""CPU: the C-ABI library loads and exports every symbol include/svae_b200.h declares
(no compute calls without a GPU)."""
import ctypes, os, re
from util import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "svae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    from svae_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    names = _declared()
    assert "svae_lds_estep_fwd" in names
    for name in names:
        assert hasattr(L, name), "missing export: %s" % name
    assert set(_lib.exported_symbols()) == set(names), "ctypes signatures out of sync with the header"
    L.svae_b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.svae_b200_version()


def test_argument_errors_without_gpu():
    from svae_b200 import _lib
    L = _lib.lib()
    # stash record = (2n + 1 + S) rows of even stride: n=3, S=1 -> 8 rows x 4 doubles
    assert L.svae_lds_estep_workspace_bytes(4, 10, 3, 1, 0) == 8 * 4 * 10 * (2 * 3 + 1 + 1) * 4
    assert L.svae_lds_estep_workspace_bytes(0, 10, 3, 1, 0) == 0


def test_no_cpu_fallback():
    """The product path must refuse CPU tensors instead of silently computing elsewhere."""
    import pytest, torch
    from svae_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(3, dtype=torch.float64))
