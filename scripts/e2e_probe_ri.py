"""Dev: piece schedules of the host pipeline in run_inference mode (H2D-bound) + raw pinned-copy bandwidth, d=10 T=200 B=4096."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from svae_b200 import synthetic as SY
from svae_b200.lds import lds_inference as K
from svae_b200.models import lds as M
n, T, B, S = 10, 200, 4096, 1
dev = torch.device("cuda")
init, pair = SY.lds_natparam(n, np.random.default_rng(0))
g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev)
nat = (tuple(g(x) for x in init), tuple(g(x) for x in pair))
hJ = (-0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64))).pin_memory()
hh = torch.randn(B, T, n, dtype=torch.float64).pin_memory()
smp = torch.empty((B, T, S, n), dtype=torch.float64).pin_memory()
dJ = torch.empty_like(hJ, device=dev)
def tm(fn, it=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it
ms = tm(lambda: dJ.copy_(hJ, non_blocking=True)); print("H2D %.1f MB: %.2f ms = %.1f GB/s" % (hJ.numel() * 8 / 1e6, ms, hJ.numel() * 8 / ms / 1e6))
ms = tm(lambda: hJ.copy_(dJ, non_blocking=True)); print("D2H %.1f MB: %.2f ms = %.1f GB/s" % (hJ.numel() * 8 / 1e6, ms, hJ.numel() * 8 / ms / 1e6))
scheds = [K.ramp_schedule(B, 256, 1.5, 1024), [256, 512, 896, 1024, 768, 448, 192], [128, 256, 512, 1024, 1024, 640, 320, 192],
          [256, 512, 1024, 1024, 768, 512], [512, 1024, 1024, 1024, 512], [384, 768, 1024, 1024, 640, 256], 8, 6, 12,
          [256, 384, 512, 640, 768, 640, 512, 256, 128]]
for ch in scheds:
    ms = tm(lambda: M.run_inference_host(None, None, (hJ, hh), S, smp, chunks=ch, device=dev, local_natparam=nat))
    print("chunks", ch, "%.2f ms  %.0f seq/s" % (ms, B / ms * 1e3), flush=True)
