"""CPU: host-side pieces of the parallel-in-time drivers that need no kernel -- the log-depth affine suffix scan that
stitches the chunk sample paths, and the choice of the default chunk length."""
import math
import numpy as np
import torch
from svae_b200.lds import lds_pit


def test_affine_suffix_scan_matches_the_sequential_recursion():
    g = torch.Generator().manual_seed(0)
    for B, C, S, n in [(2, 1, 1, 3), (1, 2, 2, 2), (3, 7, 2, 4), (1, 33, 1, 5)]:
        A = 0.7 * torch.randn(B, C, n, n, dtype=torch.float64, generator=g)
        d = torch.randn(B, C, S, n, dtype=torch.float64, generator=g)
        want = torch.zeros_like(d)
        nxt = torch.zeros(B, S, n, dtype=torch.float64)
        for c in range(C - 1, -1, -1):                    # delta_c = d_c + A_c delta_{c+1}, zero beyond the end
            nxt = d[:, c] + torch.einsum('bij,bsj->bsi', A[:, c], nxt)
            want[:, c] = nxt
        got = lds_pit._affine_suffix_scan(A, d)
        assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_default_chunk_length():
    for T in (4, 5, 50, 200, 201, 2000, 2001, 20000, 100000, 100001, 99991):
        L = lds_pit.default_chunk_len(T)
        root = math.sqrt(T - 1)
        assert 2 <= L and root / 1.6 <= L <= root * 1.6 + 1
        if T in (201, 2001, 100001, 20001):                # T-1 has a divisor near sqrt(T-1): no ragged tail
            assert (T - 1) % L == 0
    assert lds_pit.default_chunk_len(100001) == 250 and lds_pit.default_chunk_len(200) == 14


def test_ramp_schedule_of_the_host_pipeline():
    from svae_b200.lds.lds_inference import ramp_schedule
    for B in (1, 100, 255, 256, 257, 512, 4096, 5000, 100000):
        for first, growth, largest in ((256, 1.5, 1024), (128, 2.0, 2048), (512, 1.5, 1024)):
            s = ramp_schedule(B, first, growth, largest)
            assert sum(s) == B and min(s) > 0
            assert all(x <= max(largest, first) + first // 2 for x in s)          # last piece may absorb a sliver
            assert all(a <= b for a, b in zip(s[:-2], s[1:-1]))                    # non-decreasing up to the remainder
    assert ramp_schedule(4096, 256, 1.5, 1024) == [256, 384, 576, 864, 1024, 992]
