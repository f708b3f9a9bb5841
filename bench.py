#!/usr/bin/env python
"""bench.py -- LDS-SVAE local E-step throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]                  (N>1: launched by torchrun)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]

A "step" is one pass of the hot path over one minibatch of synthetic sequences: the fused
information-form Kalman filter + RTS smoother + backward sampler (one CUDA kernel launch through the
C-ABI of libsvae_b200.so) for B sequences of shape d=10, T=200 per GPU, followed by the reduction of
the global expected-sufficient-statistic accumulators over the minibatch and -- for N > 1 -- the
single NCCL all-reduce of that accumulator per SVI step.  Sequences shard across
ranks with no other collective: "scaling": "weak" (B per GPU fixed).

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the same
metric through the public API with HOST (pinned) buffers, host<->device copies inside the timed
region; `roofline` divides the algorithmic bytes of SURVEY.md §8(d) by the kernel's CUDA-event
duration; `cpu_baseline` times the reference's own compiled Cython kernels (oracle/_ref) on the host
cores.  `--impl reference` times that reference CPU path alone (rank 0 only).
"""
from __future__ import annotations
import argparse, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lds_estep_sequences_per_sec"
UNIT = "sequences/s"
D, T_LEN, BATCH, NUM_SAMPLES = 10, 200, 4096, 1


def alg_bytes_per_seq(n, T):
    """reads J_node,h_node (16Tn) + writes E_node (16Tn), samples (8Tn),
    E_init 8(n^2+n), sum E_pair 24 n^2, lognorm 8  =  8(5Tn + 4n^2 + n + 1)."""
    return 8 * (5 * T * n + 4 * n * n + n + 1)


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# --------------------------------------------------------------------------------------------------
# reference CPU arm: the reference's own Cython kernels (oracle/_ref), one sequence per call
# --------------------------------------------------------------------------------------------------

def _cpu_worker(args):
    """Runs `count` sequences of the named shape through the reference primitives; returns seconds."""
    seed, count, n, T, S = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    from oracle import ref, lds_numpy as O
    rng = np.random.default_rng(0)
    init, pair = O.rand_lds_natparam(n, rng)           # same dynamics for every worker (seed 0)
    rng = np.random.default_rng(1000 + seed)
    J, h, logZ = O.rand_node_potentials(min(count, 64), T, n, rng)
    ref.lds_estep_one(init, pair, (J[0], h[0], logZ[0]), S, seed=seed)   # warm-up / import
    t0 = time.perf_counter()
    for i in range(count):
        k = i % J.shape[0]
        ref.lds_estep_one(init, pair, (J[k], h[k], logZ[k]), S)
    return time.perf_counter() - t0


def _cpu_port_worker(args):
    seed, count, n, T, S = args
    import numpy as np
    from oracle import lds_numpy as O
    rng = np.random.default_rng(0)
    init, pair = O.rand_lds_natparam(n, rng)
    rng = np.random.default_rng(1000 + seed)
    node = O.rand_node_potentials(count, T, n, rng)
    eps = rng.standard_normal((count, T, S, n))
    t0 = time.perf_counter()
    O.natural_lds_inference_general((init, pair), node, eps)
    return time.perf_counter() - t0


def usable_cores():
    """Host threads this process can actually use: CPU affinity capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period) + 0.5)))
    except Exception:
        pass
    return n


def cpu_reference_rate(n, T, S, target_seconds=4.0, cores=None):
    """sequences/s of the reference CPU path using all host cores (one sequence stream per process)."""
    import multiprocessing as mp
    from oracle import ref
    cores = cores or usable_cores()
    kind = "reference" if ref.available() else "port"
    worker = _cpu_worker if kind == "reference" else _cpu_port_worker
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        # calibrate on one core, then size the sample for ~target_seconds of wall time on all cores
        t_one = worker((0, 8, n, T, S)) / 8
        per_proc = max(8, int(target_seconds / max(t_one, 1e-6)))
        t0 = time.perf_counter()
        times = pool.map(worker, [(i, per_proc, n, T, S) for i in range(cores)])
        wall = time.perf_counter() - t0
    total = per_proc * cores
    # rate from the slowest worker's compute time (excludes process start-up and imports)
    rate = total / max(times)
    return dict(value=rate, unit=UNIT, cores=cores, kind=kind,
                sample="%d sequences (d=%d,T=%d,S=%d) = %d per process x %d processes, %.1f s wall; "
                       "single-core %.1f seq/s" % (total, n, T, S, per_proc, cores, wall, 1.0 / t_one))


# --------------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clocks / throttle reasons with one `nvidia-smi -lms` process while the GPU is under load."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            self.lines = [l for l in out.splitlines() if l.strip()]

    def summary(self):
        rows = [[x.strip() for x in l.split(",")] for l in self.lines]
        rows = [r for r in rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in rows)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(int(r[1]) for r in rows if r[1].isdigit()),
                    reasons=reasons, samples=len(rows))


# --------------------------------------------------------------------------------------------------
# main arms
# --------------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, T, S = D, T_LEN, NUM_SAMPLES
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        # each step: a bounded sample of the workload on all host cores
        info = cpu_reference_rate(n, T, S, target_seconds=max(1.0, min(4.0, 60.0 / (args.warmup + args.steps))))
        if i >= args.warmup:
            vals.append(info["value"])
    value = sum(vals) / len(vals)
    info["value"] = value
    line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=BATCH / value * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="LDS-SVAE E-step d=%d T=%d batch=%d S=%d" % (n, T, BATCH, S),
                            note="reference Cython kernels (oracle/_ref), one sequence per call, all host cores"),
                cpu_baseline=info,
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from svae_b200 import synthetic as SY      # synthetic-problem recipe (host side, untimed); no oracle on this arm
    from svae_b200.lds import lds_inference as K
    from svae_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, T, B, S = D, T_LEN, BATCH, NUM_SAMPLES

    rng = np.random.default_rng(0)
    init, pair = SY.lds_natparam(n, rng)
    g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev)
    gi, gp = tuple(g(x) for x in init), tuple(g(x) for x in pair)
    gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)
    NSETS = 2   # rotate input sets; each set (node potentials + noise = 196 MB) exceeds the 126 MB L2
    sets = []
    for _ in range(NSETS):
        nJ = -0.5 * torch.nn.functional.softplus(torch.randn(B, T, n, dtype=torch.float64, device=dev, generator=gen))
        nh = torch.randn(B, T, n, dtype=torch.float64, device=dev, generator=gen)
        eps = torch.randn(B, T, S, n, dtype=torch.float64, device=dev, generator=gen)
        sets.append((nJ, nh, eps))

    def step(i, kev=None):
        nJ, nh, eps = sets[i % NSETS]
        if kev is not None:
            kev[0].record()
        o = K.lds_estep_forward(gi, gp, (nJ, nh), S, eps)
        if kev is not None:
            kev[1].record()
        acc = K.lds_reduce_global_stats(o)      # sum_b [E_init | E_pair | lognorm]
        if world > 1:
            dist.all_reduce(acc)                # the single collective of an SVI step
        return o, acc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with ClockSampler(local) as clk:
        # warm-up: at least W (>= 3) steps and at least ~1 s so the clock sampler sees the GPU under load
        t0 = time.perf_counter()
        i = 0
        while i < max(args.warmup, 3) or time.perf_counter() - t0 < 1.0:
            step(i); i += 1
            if i % 16 == 0:
                torch.cuda.synchronize()
        barrier()
        # ---- device-timed region: exactly K steps -------------------------------------------------
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        e0.record()
        for i in range(args.steps):
            step(i, kev[i])
        e1.record()
        barrier()
    launches = _lib.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with HOST buffers ------------------------------------
    # piece sizes of the host pipeline: small first piece so the device-to-host stream (the bound) starts early;
    # swept on a B200 in scripts/e2e_probe.py (profiles/r01_e2e_chunk_sweep.txt)
    E2E_CHUNKS = K.ramp_schedule(B, 256, 1.5, 1024)
    h_nJ = sets[0][0].cpu().pin_memory(); h_nh = sets[0][1].cpu().pin_memory()
    out_host = dict(
        samples=torch.empty((B, T, S, n), dtype=torch.float64).pin_memory(),
        E_node_xx=torch.empty((B, T, n), dtype=torch.float64).pin_memory(),
        E_node_x=torch.empty((B, T, n), dtype=torch.float64).pin_memory(),
        lognorm=torch.empty((B,), dtype=torch.float64).pin_memory())
    acc_host = torch.empty((4 * n * n + n + 1,), dtype=torch.float64).pin_memory()

    def e2e_step(i):
        acc = K.lds_estep_host((gi, gp), (h_nJ, h_nh), S, out_host, chunks=E2E_CHUNKS, generator=gen, device=dev)
        if world > 1:
            dist.all_reduce(acc)
        acc_host.copy_(acc, non_blocking=True)

    for i in range(3):
        e2e_step(i)
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(e2e_steps):
        e2e_step(i)
    f1.record()
    barrier()
    t = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    h2d = h_nJ.numel() * 8 + h_nh.numel() * 8
    d2h = sum(v.numel() * 8 for v in out_host.values()) + acc_host.numel() * 8

    # ---- secondary numbers (not the headline): training E-step (forward + VJP), HMM config ------------
    extra = {}
    if rank == 0 and not args.no_extra:
        try:
            nJ, nh, eps = sets[0]
            w = torch.randn(B, T, S, n, dtype=torch.float64, device=dev, generator=gen)
            def fb():
                o = K.lds_estep_forward(gi, gp, (nJ, nh), S, eps, save_for_bwd=True)
                K.lds_estep_backward(o, dict(samples=w, E_node_xx=nJ, E_node_x=nh, lognorm=-torch.ones(B, dtype=torch.float64, device=dev)))
            for _ in range(2):
                fb()
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fb()
            b_.record(); torch.cuda.synchronize()
            extra["lds_estep_fwd_plus_vjp_seq_per_s"] = B / (a.elapsed_time(b_) / 5 * 1e-3)
            from svae_b200.hmm import hmm_inference as HK
            lp, lA = SY.hmm_natparam(32, rng)
            hnode = torch.randn(2048, 500, 32, dtype=torch.float64, device=dev, generator=gen)
            nat = (g(lp), g(lA), hnode)
            for _ in range(2):
                HK.hmm_estep(nat)
            torch.cuda.synchronize()
            a.record()
            for _ in range(5):
                HK.hmm_estep(nat)
            b_.record(); torch.cuda.synchronize()
            extra["hmm_estep_K32_T500_B2048_seq_per_s"] = 2048 / (a.elapsed_time(b_) / 5 * 1e-3)
            # the other LDS shapes BASELINE.json lists (parity-test cases; reported for context only)
            for (n2, T2, B2) in ((8, 50, 256), (16, 200, 4096)):
                i2, p2 = SY.lds_natparam(n2, rng)
                gi2, gp2 = tuple(g(x) for x in i2), tuple(g(x) for x in p2)
                J2 = -0.5 * torch.nn.functional.softplus(torch.randn(B2, T2, n2, dtype=torch.float64, device=dev, generator=gen))
                h2 = torch.randn(B2, T2, n2, dtype=torch.float64, device=dev, generator=gen)
                e2 = torch.randn(B2, T2, 1, n2, dtype=torch.float64, device=dev, generator=gen)
                for _ in range(2):
                    K.lds_estep_forward(gi2, gp2, (J2, h2), 1, e2)
                torch.cuda.synchronize()
                a.record()
                for _ in range(5):
                    K.lds_estep_forward(gi2, gp2, (J2, h2), 1, e2)
                b_.record(); torch.cuda.synchronize()
                extra["lds_estep_d%d_T%d_B%d_seq_per_s" % (n2, T2, B2)] = B2 / (a.elapsed_time(b_) / 5 * 1e-3)
            # small-batch / long-T regime: one sequence of 20000 steps, sequential vs
            # parallel-in-time driver (same outputs, tests/test_lds_pit_gpu.py)
            from svae_b200.lds import lds_pit
            T3 = 20000
            J3 = -0.5 * torch.nn.functional.softplus(torch.randn(1, T3, n, dtype=torch.float64, device=dev, generator=gen))
            h3 = torch.randn(1, T3, n, dtype=torch.float64, device=dev, generator=gen)
            e3 = torch.randn(1, T3, 1, n, dtype=torch.float64, device=dev, generator=gen)
            for name, fn in (("sequential", lambda: K.natural_lds_inference_general((gi, gp), (J3, h3), 1, e3)),
                             ("parallel_in_time", lambda: lds_pit.natural_lds_inference_parallel((gi, gp), (J3, h3), 1, e3))):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                a.record()
                for _ in range(5):
                    fn()
                b_.record(); torch.cuda.synchronize()
                extra["lds_estep_d%d_T%d_B1_%s_ms" % (n, T3, name)] = a.elapsed_time(b_) / 5
            from svae_b200.hmm import hmm_pit
            hnode1 = torch.randn(1, 100000, 32, dtype=torch.float64, device=dev, generator=gen)
            nat1 = (g(lp), g(lA), hnode1)
            for name, fn in (("sequential", lambda: HK.hmm_estep(nat1)), ("parallel_in_time", lambda: hmm_pit.hmm_estep_parallel(nat1))):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                a.record()
                for _ in range(5):
                    fn()
                b_.record(); torch.cuda.synchronize()
                extra["hmm_estep_K32_T100000_B1_%s_ms" % name] = a.elapsed_time(b_) / 5
        except Exception as e:  # pragma: no cover
            extra["error"] = repr(e)

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        ab = alg_bytes_per_seq(n, T) * B + 8 * T * S * n * B     # + injected noise read
        achieved = ab / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("lds_estep_fwd_d10_T200_B4096")
            except Exception:
                traffic = None
        try:
            if args.no_cpu_baseline:
                raise RuntimeError("skipped (--no-cpu-baseline)")
            cpu = cpu_reference_rate(n, T, S, target_seconds=4.0)
        except Exception as e:  # pragma: no cover
            cpu = dict(value=None, unit=UNIT, cores=0, kind="unavailable", sample=repr(e))
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
            data="synthetic",
            config=dict(workload="LDS-SVAE E-step d=%d T=%d batch=%d per GPU S=%d" % (n, T, B, S),
                        global_batch=world * B, parallelism="dp%d" % world,
                        l2="inputs per step (196 MB node potentials+noise, 1.4 GB stash) exceed the 126 MB L2; "
                           "%d input sets rotated" % NSETS,
                        step="fused filter+smoother+sampler kernel, 2-stage minibatch reduction of the global statistics"
                             + (", NCCL all-reduce of the %d-double accumulator" % (4 * n * n + n + 1) if world > 1 else "")),
            e2e=dict(value=world * B / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     ms_per_step=e2e_ms, note="pinned host buffers; pieces of %s sequences pipelined (one copy stream per direction, kernels on 3 streams), persistent device buffers; noise drawn on device" % (E2E_CHUNKS,)),
            gpu_launches=int(launches),
            clocks=clk.summary(),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=traffic, peak_source=peak_src, kernel="lds_estep_fwd_kernel<10,16,false>",
                          kernel_ms=kern_ms, algorithmic_bytes_per_launch=ab,
                          note="fp64 ALU / shared-memory bound at ~30 flop/B, not HBM bound (SURVEY.md 8d, DESIGN.md 3)",
                          # the second roofline SURVEY.md 8d asks for: algorithmic flops T(12n^3+10n^2) per sequence
                          # (the REFERENCE's operation count; this kernel's own formulation needs ~30% fewer) over the
                          # fp64 FMA peak 148 SMs x 64 FMA/clk x 2 x 1.965 GHz
                          fp64=dict(achieved_tflops=T * (12 * n ** 3 + 10 * n ** 2) * B / (kern_ms * 1e-3) / 1e12,
                                    peak_tflops=148 * 64 * 2 * 1.965e9 / 1e12,
                                    frac=T * (12 * n ** 3 + 10 * n ** 2) * B / (kern_ms * 1e-3) / (148 * 64 * 2 * 1.965e9))),
            cpu_baseline=cpu, extra=extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-core baseline (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary (fwd+VJP, HMM) timings")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
