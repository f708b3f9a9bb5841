"""Dev helper: time build variants of libsvae_b200 against the product library in ONE GPU call and check that every
variant reproduces the product library's outputs.

    python scripts/variant_bench.py [tag ...]        # tags of svae_b200/libsvae_b200_<tag>.so; "" = the product library

Each library runs in its own process (SVAE_B200_LIB is read once at import).  The product library runs first and leaves
its outputs under /tmp; a variant's outputs are compared against them (max abs difference relative to max abs value).
Timing: CUDA events, 3 warm-up + 20 launches, inputs resident (larger than L2 at the LDS shapes).
"""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(tag):
    import numpy as np, torch
    from svae_b200 import synthetic as SY
    from svae_b200.lds import lds_inference as K
    from svae_b200.hmm import hmm_inference as HK
    from svae_b200 import _lib
    dev = torch.device("cuda", 0)
    ref_dir = "/tmp/variant_ref"
    os.makedirs(ref_dir, exist_ok=True)

    def timeit(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def compare(name, outs):
        path = os.path.join(ref_dir, name + ".pt")
        cpu = {k: v.detach().cpu() for k, v in outs.items() if torch.is_tensor(v)}
        if not tag:
            torch.save(cpu, path)
            return 0.0
        if not os.path.exists(path):
            return float("nan")
        ref = torch.load(path)
        worst = 0.0
        for k, v in cpu.items():
            r = ref[k]
            if v.dtype.is_floating_point:
                worst = max(worst, float((v - r).abs().max() / r.abs().max().clamp_min(1e-300)))
            else:
                worst = max(worst, float((v != r).any()))
        return worst

    def lds_case(n, T, B, S, save=False):
        g = torch.Generator(device=dev).manual_seed(1234 + n)
        rng = np.random.default_rng(n)
        init, pair = SY.lds_natparam(n, rng)
        t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev)
        gi, gp = tuple(t(x) for x in init), tuple(t(x) for x in pair)
        nJ = -0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64, device=dev, generator=g))
        nh = torch.randn(B, T, n, dtype=torch.float64, device=dev, generator=g)
        eps = torch.randn(B, T, S, n, dtype=torch.float64, device=dev, generator=g) if S else None
        bufs = {}
        o = K.lds_estep_forward(gi, gp, (nJ, nh), S, eps, save_for_bwd=save)
        keys = ["lognorm", "E_node_xx", "E_node_x", "samples", "E_init_xx", "E_init_x", "E_pair_xx", "E_pair_xxn", "E_pair_xnxn"]
        small = {k: (o[k][:64] if o[k].dim() > 1 and o[k].shape[0] > 64 else o[k]) for k in keys if o.get(k) is not None}
        small["sum_all"] = torch.stack([o[k].sum() for k in keys if o.get(k) is not None])
        err = compare("lds_n%d_T%d_B%d_S%d_%d" % (n, T, B, S, save), small)
        bufs = {k: o[k] for k in keys if o.get(k) is not None}
        bufs["workspace"] = o.get("workspace")
        bufs["status"] = o["status"]
        if bufs["workspace"] is None:
            L = _lib.lib()
            bufs["workspace"] = torch.empty(L.svae_lds_estep_workspace_bytes(B, T, n, S, 0), dtype=torch.uint8, device=dev)
        ms = timeit(lambda: K.lds_estep_forward(gi, gp, (nJ, nh), S, eps, save_for_bwd=save, buffers=bufs))
        print("[%s] LDS n=%d T=%d B=%d S=%d save=%d: %.4f ms  %.3e seq/s   vs product outputs: %.2e"
              % (tag or "product", n, T, B, S, save, ms, B / ms * 1e3, err), flush=True)

    def vjp_case(n, T, B, S):
        g = torch.Generator(device=dev).manual_seed(4321 + n)
        rng = np.random.default_rng(n)
        init, pair = SY.lds_natparam(n, rng)
        t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev)
        gi, gp = tuple(t(x) for x in init), tuple(t(x) for x in pair)
        r = lambda *sh: torch.randn(*sh, dtype=torch.float64, device=dev, generator=g)
        nJ, nh, eps, w = -0.5 * torch.nn.functional.softplus(r(B, T, n)), r(B, T, n), r(B, T, S, n), r(B, T, S, n)
        gl = -torch.ones(B, dtype=torch.float64, device=dev)
        fwd = lambda: K.lds_estep_forward(gi, gp, (nJ, nh), S, eps, save_for_bwd=True)
        bwd = lambda o: K.lds_estep_backward(o, dict(samples=w, E_node_xx=nJ, E_node_x=nh, lognorm=gl))
        gJ, gh = bwd(fwd())
        err = compare("vjp_n%d_T%d_B%d_S%d" % (n, T, B, S), dict(gJ=gJ[:64], gh=gh[:64], s=torch.stack([gJ.sum(), gh.sum()])))
        for _ in range(2):
            bwd(fwd())
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        iters = 6
        for _ in range(iters):
            ev[0].record(); o = fwd(); ev[1].record(); bwd(o); ev[2].record(); torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
        print("[%s] VJP n=%d T=%d B=%d S=%d: fwd(save) %.4f ms  bwd %.4f ms  %.3e seq/s   vs product outputs: %.2e"
              % (tag or "product", n, T, B, S, tf / iters, tb / iters, B / ((tf + tb) / iters) * 1e3, err), flush=True)

    def hmm_case(Kk, T, B):
        g = torch.Generator(device=dev).manual_seed(77 + Kk)
        rng = np.random.default_rng(Kk)
        lp, lA = SY.hmm_natparam(Kk, rng)
        t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev)
        node = torch.randn(B, T, Kk, dtype=torch.float64, device=dev, generator=g)
        nat = (t(lp), t(lA), node)
        logZ, (E0, Et, Ez) = HK.hmm_estep(nat)
        err = compare("hmm_K%d_T%d_B%d" % (Kk, T, B), dict(logZ=logZ, E0=E0, Et=Et, Ez=Ez[:64]))
        ms = timeit(lambda: HK.hmm_estep(nat))
        print("[%s] HMM K=%d T=%d B=%d: %.4f ms  %.3e seq/s   vs product outputs: %.2e"
              % (tag or "product", Kk, T, B, ms, B / ms * 1e3, err), flush=True)

    cases = os.environ.get("VARIANT_CASES", "lds10,lds16,lds10s0,hmm32,hmm20,hmmT").split(",")
    if "lds10" in cases: lds_case(10, 200, 4096, 1)
    if "lds10s0" in cases: lds_case(10, 200, 4096, 0)
    if "lds16" in cases: lds_case(16, 200, 4096, 1)
    if "lds10small" in cases: lds_case(10, 7, 5, 1)
    if "vjp10" in cases: vjp_case(10, 200, 4096, 1)
    if "vjp16" in cases: vjp_case(16, 200, 4096, 1)
    if "vjp12" in cases: vjp_case(12, 200, 4096, 1)
    if "vjp14" in cases: vjp_case(14, 200, 4096, 1)
    if "hmm32" in cases: hmm_case(32, 500, 2048)
    if "hmmbig" in cases: hmm_case(32, 500, 16384)
    if "hmm24" in cases: hmm_case(24, 500, 2048)
    if "hmm20" in cases: hmm_case(20, 500, 2048)
    if "hmmT" in cases:
        for T in (1, 2, 3, 6, 7, 11):
            hmm_case(32, T, 5)
            hmm_case(17, T, 3)


def main():
    tags = sys.argv[1:] or [""]
    for spec in tags:          # spec = tag[:case,case...][@ENV=VALUE...]   (tag "" or "product" = the product library)
        env = dict(os.environ)
        parts = spec.split("@")
        for kv in parts[1:]:
            k, v = kv.split("=", 1)
            env[k] = v
        tag, _, cases = parts[0].partition(":")
        if cases:
            env["VARIANT_CASES"] = cases
        lib = "" if tag in ("", "product") else tag
        if lib:
            env["SVAE_B200_LIB"] = os.path.join(ROOT, "svae_b200", "libsvae_b200_%s.so" % lib)
        tag = spec if len(parts) > 1 else lib
        env["VARIANT_TAG"] = tag
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=env)
        if r.returncode != 0:
            print("[%s] worker failed: rc %d" % (tag or "product", r.returncode), flush=True)


if __name__ == "__main__":
    if "--worker" in sys.argv:
        worker(os.environ.get("VARIANT_TAG", ""))
    else:
        main()
