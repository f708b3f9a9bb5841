"""Dev helper: time the fused LDS E-step kernel at a given shape (CUDA events, inputs resident)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lds_numpy as O
from svae_b200.lds import lds_inference as K

def main():
    n, T, B, S = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (10, 200, 4096, 1))]
    iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    rng = np.random.default_rng(0)
    init, pair = O.rand_lds_natparam(n, rng)
    g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device="cuda")
    gi, gp = tuple(g(x) for x in init), tuple(g(x) for x in pair)
    nJ = -0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64, device="cuda"))
    nh = torch.randn(B, T, n, dtype=torch.float64, device="cuda")
    eps = torch.randn(B, T, S, n, dtype=torch.float64, device="cuda") if S else None
    for _ in range(3):
        K.lds_estep_forward(gi, gp, (nJ, nh), S, eps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        K.lds_estep_forward(gi, gp, (nJ, nh), S, eps)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    abytes = 8 * (5 * T * n + 4 * n * n + n + 1) * B
    print("n=%d T=%d B=%d S=%d: %.3f ms/launch  %.3e seq/s  alg %.1f GB/s" % (n, T, B, S, ms, B / ms * 1e3, abytes / ms / 1e6))

if __name__ == "__main__":
    main()
