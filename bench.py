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
import argparse, json, os, subprocess, sys, time

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
    """One process of the reference CPU arm: runs `count` units of work of the named kind, returns seconds.
    kinds: lds_fwd / lds_fwd_bwd / hmm (the reference's own compiled Cython, oracle/_ref), *_port (NumPy restatement
    when oracle/_ref is not built), slds_port / gmm_port (oracle restatements pinned to the reference's Python by
    tests/test_reference_py_fixtures.py -- the reference's Python layer itself cannot run on the GPU box)."""
    kind, seed, count, shape = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    rng0 = np.random.default_rng(0)
    rng = np.random.default_rng(1000 + seed)
    if kind in ("lds_fwd", "lds_fwd_bwd"):
        from oracle import ref, lds_numpy as O
        n, T, S = shape
        init, pair = O.rand_lds_natparam(n, rng0)           # same dynamics for every worker (seed 0)
        J, h, logZ = O.rand_node_potentials(min(count, 64), T, n, rng)
        if kind == "lds_fwd":
            run = lambda k: ref.lds_estep_one(init, pair, (J[k], h[k], logZ[k]), S)
        else:
            gs = rng.standard_normal((T, S, n))
            gst = ((np.zeros((n, n)), np.zeros(n), 0., 0.), (np.zeros((n, n)),) * 3 + (0.,), (J[0], h[0], np.zeros(T)))
            run = lambda k: ref.lds_vjp_one(init, pair, (J[k], h[k], logZ[k]), gs, gst, -1.0, num_samples=S)
        run(0)                                              # warm-up / import
        t0 = time.perf_counter()
        for i in range(count):
            run(i % J.shape[0])
        return time.perf_counter() - t0
    if kind == "lds_fwd_port":
        from oracle import lds_numpy as O
        n, T, S = shape
        init, pair = O.rand_lds_natparam(n, rng0)
        node = O.rand_node_potentials(count, T, n, rng)
        eps = rng.standard_normal((count, T, S, n))
        t0 = time.perf_counter()
        O.natural_lds_inference_general((init, pair), node, eps)
        return time.perf_counter() - t0
    if kind == "hmm":
        from oracle import ref, hmm_numpy as H
        K, T = shape
        lp, lA = H.rand_hmm_natparam(K, rng0)
        node = rng.standard_normal((min(count, 16), T, K))
        ref.hmm_estep_slow((lp, lA, node[0]))
        t0 = time.perf_counter()
        for i in range(count):
            ref.hmm_estep_slow((lp, lA, node[i % node.shape[0]]))   # hmm_logZ + hmm_logZ_grad (hmm_inference.py:65)
        return time.perf_counter() - t0
    if kind == "slds_port":
        from oracle import slds_numpy as SO, lds_numpy as O
        from svae_b200 import synthetic as SY
        K, n, T = shape
        glob = SY.slds_global_natparam(K, n, rng0)
        t0 = time.perf_counter()
        for i in range(count):
            node = O.rand_node_potentials(1, T, n, rng)[:2]
            hmm, lds_all, nat, _, _ = SO.optimize_local_meanfield(glob, node, rng.standard_normal((1, T, 1, n)))
            init3 = (nat[0][0][0], nat[0][1][0], nat[0][2][0] + nat[0][3][0])
            O.natural_lds_inference_general((init3, tuple(p[0] for p in nat[1])), node + (np.zeros((1, T)),),
                                            rng.standard_normal((1, T, 1, n)))
        return time.perf_counter() - t0
    if kind == "gmm_port":
        from oracle import gmm_numpy as G
        from svae_b200 import synthetic as SY
        K, N, T = shape
        glob = SY.gmm_global_natparam(K, N, rng0)
        t0 = time.perf_counter()
        for i in range(count):
            node = (-0.5 * np.log1p(np.exp(rng.standard_normal((T, N)))), 2.0 * rng.standard_normal((T, N)))
            init = rng.random((T, K)); init /= init.sum(-1, keepdims=True)
            G.local_meanfield(glob, node, init)
        return time.perf_counter() - t0
    raise ValueError(kind)


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


class CpuArm:
    """The reference CPU path on all usable host cores: one spawn pool, one stream of work units per process."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or usable_cores()
        self.pool = mp.get_context("spawn").Pool(self.cores)

    def close(self):
        self.pool.close(); self.pool.join()

    def rate(self, kind, shape, unit_name, target_seconds=4.0, calib=4):
        """units/s; a bounded sample sized for ~target_seconds of wall time."""
        from oracle import ref
        label = "reference"
        if kind in ("lds_fwd", "lds_fwd_bwd", "hmm") and not ref.available():
            if kind != "lds_fwd":
                raise RuntimeError("oracle/_ref is not built")
            kind, label = "lds_fwd_port", "port"
        if kind.endswith("_port"):
            label = "port"
        t_one = self.pool.apply(_cpu_worker, ((kind, 0, calib, shape),)) / calib
        per_proc = max(2, int(target_seconds / max(t_one, 1e-6)))
        t0 = time.perf_counter()
        times = self.pool.map(_cpu_worker, [(kind, i, per_proc, shape) for i in range(self.cores)])
        wall = time.perf_counter() - t0
        total = per_proc * self.cores
        return dict(value=total / max(times), unit=unit_name, cores=self.cores, kind=label,
                    sample="%d units %s%s = %d per process x %d processes, %.1f s wall; single-core %.2f /s"
                           % (total, kind, tuple(shape), per_proc, self.cores, wall, 1.0 / t_one))


def cpu_reference_rate(n, T, S, target_seconds=4.0, cores=None):
    """sequences/s of the reference CPU E-step (filter + smoother + sampler) using all host cores."""
    arm = CpuArm(cores)
    try:
        return arm.rate("lds_fwd", (n, T, S), UNIT, target_seconds, calib=8)
    finally:
        arm.close()


def bind_to_gpu_numa_node(index):
    """Pin this rank's threads (and so its first-touch pinned host buffers) to the NUMA node of its GPU: ranks on the
    far socket otherwise push every PCIe copy across the inter-socket link.  Returns a short description."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        bus = out.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return "numa: node %d has no allowed cpu" % node
        os.sched_setaffinity(0, allowed)
        return "numa: gpu %d (%s) -> node %d, %d cpus" % (index, bus, node, len(allowed))
    except Exception as e:  # pragma: no cover
        return "numa: not bound (%r)" % (e,)


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

WORKLOAD = "LDS-SVAE E-step d=%d T=%d batch=%d per GPU S=%d" % (D, T_LEN, BATCH, NUM_SAMPLES)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, T, S = D, T_LEN, NUM_SAMPLES
    vals = []
    info = None
    arm = CpuArm()
    try:
        for i in range(args.warmup + args.steps):
            # each step: a bounded sample of the workload on all host cores
            info = arm.rate("lds_fwd", (n, T, S), UNIT, calib=8,
                            target_seconds=max(1.0, min(4.0, 60.0 / (args.warmup + args.steps))))
            if i >= args.warmup:
                vals.append(info["value"])
    finally:
        arm.close()
    value = sum(vals) / len(vals)
    info["value"] = value
    line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=BATCH / value * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD,
                            note="reference Cython kernels (oracle/_ref), one sequence per call, all host cores"),
                cpu_baseline=info,
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def _time_gpu(torch, fn, iters=5, warm=2):
    """ms per call (CUDA events on the current stream)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def _roof(bytes_per_launch, ms, peak, flops=None):
    r = dict(bound="hbm", achieved=bytes_per_launch / (ms * 1e-3) / 1e9, peak=peak, unit="GB/s",
             frac=bytes_per_launch / (ms * 1e-3) / 1e9 / peak, kernel_ms=ms, algorithmic_bytes_per_launch=bytes_per_launch)
    if flops is not None:
        pk = 148 * 64 * 2 * 1.965e9
        r["fp64"] = dict(achieved_tflops=flops / (ms * 1e-3) / 1e12, peak_tflops=pk / 1e12, frac=flops / (ms * 1e-3) / pk)
    return r


def run_parity(torch, K, ref_natparam, dev, n, T, S, nseq=8):
    """the GPU path against the compiled reference (oracle/_ref; the NumPy port if it is not
    built) on an `nseq`-sequence sub-batch IN THIS RUN, same inputs, same noise.  The oracle is the checker only."""
    import numpy as np
    from oracle import ref, lds_numpy as O
    init, pair = ref_natparam
    rng = np.random.default_rng(7)
    J, h, logZ = O.rand_node_potentials(nseq, T, n, rng)
    want = {k: [] for k in ("samples", "lognorm", "E_node_xx", "E_node_x", "E_init_xx", "E_pair_xx", "E_pair_xxn", "E_pair_xnxn")}
    if ref.available():
        kind, eps = "reference", []
        for b in range(nseq):
            o = ref.lds_estep_one(init, pair, (J[b], h[b], logZ[b]), S, seed=100 + b)
            eps.append(o["eps"])
            for k, v in (("samples", o["samples"]), ("lognorm", o["lognorm"]), ("E_node_xx", o["E_node"][0]),
                         ("E_node_x", o["E_node"][1]), ("E_init_xx", o["E_init"][0]), ("E_pair_xx", o["E_pair"][0]),
                         ("E_pair_xxn", o["E_pair"][1]), ("E_pair_xnxn", o["E_pair"][2])):
                want[k].append(np.asarray(v))
        eps = np.stack(eps)
        want = {k: np.stack(v) for k, v in want.items()}
    else:
        kind = "port"
        eps = rng.standard_normal((nseq, T, S, n))
        smp, (Ei, Ep, En), ln = O.natural_lds_inference_general((init, pair), (J, h, logZ), eps)
        want = dict(samples=smp, lognorm=ln, E_node_xx=En[0], E_node_x=En[1], E_init_xx=Ei[0], E_pair_xx=Ep[0],
                    E_pair_xxn=Ep[1], E_pair_xnxn=Ep[2])
    g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev)
    o = K.lds_estep_forward(tuple(g(x) for x in init), tuple(g(x) for x in pair), (g(J), g(h), g(logZ)), S, g(eps))
    errs = {}
    for k, w in want.items():
        got = o[k].cpu().numpy()
        errs[k] = float(np.abs(got - w).max() / max(np.abs(w).max(), 1e-300))
    tol = 1e-5
    worst = max(errs.values())
    assert worst <= tol, ("in-run parity against the %s failed" % kind, errs)
    return dict(checker=kind, sequences=nseq, max_rel_err=worst, tolerance=tol, per_output=errs)


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
    numa = bind_to_gpu_numa_node(local)        # before any pinned allocation
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a mismatched collective must fail within minutes, not hold the box for NCCL's default 10
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    n, T, B, S = D, T_LEN, BATCH, NUM_SAMPLES

    rng = np.random.default_rng(0)
    init, pair = SY.lds_natparam(n, rng)
    g = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev)
    gi, gp = tuple(g(x) for x in init), tuple(g(x) for x in pair)
    gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)
    rnd = lambda *s: torch.randn(*s, dtype=torch.float64, device=dev, generator=gen)
    softplus = torch.nn.functional.softplus
    NSETS = 2   # rotate input sets; each set (node potentials + noise = 196 MB) exceeds the 126 MB L2
    sets = []
    for _ in range(NSETS):
        sets.append((-0.5 * softplus(rnd(B, T, n)), rnd(B, T, n), rnd(B, T, S, n)))

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
        # warm-up: at least W (>= 3) steps and at least ~1 s so the clock sampler sees the GPU under load.  Every rank must
        # issue the SAME number of steps (each carries a collective): blocks of 32 steps, and the decision to go on is itself
        # all-reduced (a rank-local clock check would let ranks disagree by a step and dead-lock the all-reduce sequence).
        t0 = time.perf_counter()
        i = 0
        for _ in range(max(args.warmup, 3)):
            step(i); i += 1
        while True:
            torch.cuda.synchronize()
            more = torch.tensor([1.0 if time.perf_counter() - t0 < 1.0 else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(more, op=dist.ReduceOp.MAX)
            if float(more.item()) == 0.0:
                break
            for _ in range(32):
                step(i); i += 1
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
    RI_CHUNKS = K.trapezoid_schedule(B)      # host-to-device bound: scripts/e2e_probe_ri.py
    h_nJ = sets[0][0].cpu().pin_memory(); h_nh = sets[0][1].cpu().pin_memory()
    out_host = dict(
        samples=torch.empty((B, T, S, n), dtype=torch.float64).pin_memory(),
        E_node_xx=torch.empty((B, T, n), dtype=torch.float64).pin_memory(),
        E_node_x=torch.empty((B, T, n), dtype=torch.float64).pin_memory(),
        lognorm=torch.empty((B,), dtype=torch.float64).pin_memory())
    acc_host = torch.empty((4 * n * n + n + 1,), dtype=torch.float64).pin_memory()
    for v in out_host.values():
        v.zero_()                                # first touch on this rank's NUMA node

    from svae_b200.models import lds as LDSM

    def e2e_step_all(i):
        """primitive level: every per-sequence output of natural_lds_inference_general goes back to the host"""
        acc = K.lds_estep_host((gi, gp), (h_nJ, h_nh), S, out_host, chunks=E2E_CHUNKS, generator=gen, device=dev)
        if world > 1:
            dist.all_reduce(acc)
        acc_host.copy_(acc, non_blocking=True)

    ri_host = torch.empty(((n + 2) * (n + 2) + 3 * n * n + 2,), dtype=torch.float64).pin_memory()   # packed NIW stats | MNIW stats | local_kl

    def e2e_step_ri(i):
        """model level (what make_gradfun calls, svae.py:21-22): lds.run_inference for a host-resident minibatch -- samples,
        the summed global statistics and local_kl come back; the node statistics feed local_kl on the device"""
        _, stats, _, local_kl = LDSM.run_inference_host(None, None, (h_nJ, h_nh), S, out_host["samples"], chunks=RI_CHUNKS,
                                                        generator=gen, device=dev, local_natparam=(gi, gp))
        acc = torch.cat([stats[0].reshape(-1)] + [x.reshape(-1) for x in stats[1]] + [local_kl.reshape(1)])
        if world > 1:
            dist.all_reduce(acc)            # the one collective of an SVI step: global statistics + local KL
        ri_host.copy_(acc, non_blocking=True)

    def time_e2e(fn):
        for i in range(3):
            fn(i)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(e2e_steps):
            fn(i)
        f1.record()
        barrier()
        tt = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / e2e_steps

    e2e_steps = max(3, min(args.steps, 20))
    e2e_all_ms = time_e2e(e2e_step_all)
    e2e_ms = time_e2e(e2e_step_ri)
    h2d = h_nJ.numel() * 8 + h_nh.numel() * 8
    d2h_all = sum(v.numel() * 8 for v in out_host.values()) + acc_host.numel() * 8
    d2h = out_host["samples"].numel() * 8 + ri_host.numel() * 8
    del h_nJ, h_nh, out_host

    peak, peak_src = measured_peak_hbm()
    extra = {"numa": numa,
             "e2e_all_outputs": dict(value=world * B / (e2e_all_ms * 1e-3), unit=UNIT, ms_per_step=e2e_all_ms, h2d_bytes_per_step=h2d,
                                     d2h_bytes_per_step=d2h_all,
                                     note="primitive level (natural_lds_inference_general): samples AND the per-sequence node "
                                          "statistics E_node_xx, E_node_x, lognorm copied back; floor = that %.0f MB device-to-host "
                                          "copy (~4.0 ms at the ~50 GB/s PCIe 5 x16 delivers)" % (d2h_all / 1e6))}

    # ---- strong scaling (BASELINE.json configs[3]): a GLOBAL batch of 4096 sequences sharded over the ranks, same
    #      step (kernel + minibatch reduction + the one all-reduce); measured at every N, max over ranks ----------
    if not args.no_extra:
        strong = {}
        for (n2, T2, Bg) in ((16, 200, 4096), (10, 200, 4096)):
            Bl = Bg // world
            i2, p2 = SY.lds_natparam(n2, np.random.default_rng(n2))
            gi2, gp2 = tuple(g(x) for x in i2), tuple(g(x) for x in p2)
            J2, h2, e2 = -0.5 * softplus(rnd(Bl, T2, n2)), rnd(Bl, T2, n2), rnd(Bl, T2, 1, n2)

            def sstep():
                o = K.lds_estep_forward(gi2, gp2, (J2, h2), 1, e2)
                acc = K.lds_reduce_global_stats(o)
                if world > 1:
                    dist.all_reduce(acc)
            barrier()
            ms = _time_gpu(torch, sstep, iters=10, warm=3)
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
            strong["lds_d%d_T%d_global_batch_%d" % (n2, T2, Bg)] = dict(
                n_gpus=world, sequences_per_gpu=Bl, ms_per_step=ms, value=Bg / (ms * 1e-3), unit=UNIT, scaling="strong")
            del J2, h2, e2
        extra["strong_scaling"] = strong

    # ---- the other BASELINE.json configs (rank 0, N = 1): each with its roofline and the reference CPU rate -----------
    cpu = dict(value=None, unit=UNIT, cores=0, kind="unavailable", sample="skipped")
    if rank == 0:
        arm = None
        try:
            if args.no_cpu_baseline:
                raise RuntimeError("skipped (--no-cpu-baseline)")
            arm = CpuArm()
            cpu = arm.rate("lds_fwd", (n, T, S), UNIT, target_seconds=4.0, calib=8)
        except Exception as e:  # pragma: no cover
            cpu = dict(value=None, unit=UNIT, cores=0, kind="unavailable", sample=repr(e))

        def cpu_rate(kind, shape, unit, secs=2.0, calib=2):
            if arm is None:
                return None
            try:
                return arm.rate(kind, shape, unit, target_seconds=secs, calib=calib)
            except Exception as e:  # pragma: no cover
                return dict(value=None, kind="unavailable", sample=repr(e))

        if not args.no_cpu_baseline:
            try:
                extra["parity"] = run_parity(torch, K, (init, pair), dev, n, T, S)
            except AssertionError:
                raise
            except Exception as e:  # pragma: no cover
                extra["parity"] = dict(error=repr(e))
        if world == 1 and not args.no_extra:
            cfg = {}
            try:
                # LDS d=8 T=50 B=256 and d=16 T=200 B=4096 (forward E-step)
                for (n2, T2, B2) in ((8, 50, 256), (16, 200, 4096)):
                    i2, p2 = SY.lds_natparam(n2, np.random.default_rng(n2))
                    gi2, gp2 = tuple(g(x) for x in i2), tuple(g(x) for x in p2)
                    J2, h2, e2 = -0.5 * softplus(rnd(B2, T2, n2)), rnd(B2, T2, n2), rnd(B2, T2, 1, n2)
                    ms = _time_gpu(torch, lambda: K.lds_estep_forward(gi2, gp2, (J2, h2), 1, e2), iters=10)
                    cfg["lds_d%d_T%d_B%d" % (n2, T2, B2)] = dict(
                        value=B2 / (ms * 1e-3), unit=UNIT, ms=ms,
                        roofline=_roof((alg_bytes_per_seq(n2, T2) + 8 * T2 * n2) * B2, ms, peak,
                                       flops=T2 * (12 * n2 ** 3 + 10 * n2 ** 2) * B2),
                        cpu_baseline=cpu_rate("lds_fwd", (n2, T2, 1), UNIT))
                    del J2, h2, e2
                # training E-step: forward (saving for the VJP) + reverse-mode kernel
                nJ, nh, eps = sets[0]
                w = rnd(B, T, S, n)
                gl = -torch.ones(B, dtype=torch.float64, device=dev)

                def fb():
                    o = K.lds_estep_forward(gi, gp, (nJ, nh), S, eps, save_for_bwd=True)
                    K.lds_estep_backward(o, dict(samples=w, E_node_xx=nJ, E_node_x=nh, lognorm=gl))
                ms = _time_gpu(torch, fb, iters=5)
                cfg["lds_fwd_plus_vjp_d%d_T%d_B%d" % (n, T, B)] = dict(
                    value=B / (ms * 1e-3), unit=UNIT, ms=ms,
                    roofline=_roof((alg_bytes_per_seq(n, T) + 8 * T * n + 40 * T * n) * B, ms, peak),   # + cotangents in, gradients out
                    cpu_baseline=cpu_rate("lds_fwd_bwd", (n, T, S), UNIT))
                del w
                # HMM K=32 T=500 B=2048
                from svae_b200.hmm import hmm_inference as HK
                Kh, Th, Bh = 32, 500, 2048
                lp, lA = SY.hmm_natparam(Kh, rng)
                nat = (g(lp), g(lA), rnd(Bh, Th, Kh))
                ms = _time_gpu(torch, lambda: HK.hmm_estep(nat), iters=5)
                cfg["hmm_K%d_T%d_B%d" % (Kh, Th, Bh)] = dict(
                    value=Bh / (ms * 1e-3), unit=UNIT, ms=ms,
                    roofline=_roof(8 * (2 * Th * Kh + Kh * Kh + Kh + 1) * Bh, ms, peak),
                    kernels="hmm_tc_sweep4_kernel + hmm_tc_trans_kernel (fp64 tensor-core instruction mma.m8n8k4, 8 sequences "
                            "per fragment row block; workspace 2 B T K doubles allocated inside the timed call)",
                    cpu_baseline=cpu_rate("hmm", (Kh, Th), UNIT))
                del nat
                # SLDS K=8 d=10 T=200 B=1024: run_inference (mean-field fixed point + differentiable evaluation)
                from svae_b200.models import slds_svae as SL
                Ks, Bs = 8, 1024
                tg = lambda x: g(x)
                (dn, mn), lds_np = SY.slds_global_natparam(Ks, n, np.random.default_rng(0))
                sglob = ((tg(dn), tg(mn)), [(tg(a), tuple(tg(x) for x in b)) for a, b in lds_np])
                sJ, sh = -0.5 * softplus(rnd(Bs, T, n)), rnd(Bs, T, n)
                ms = _time_gpu(torch, lambda: SL.run_inference(sglob, sglob, (sJ, sh), 1, generator=gen), iters=3, warm=1)
                cfg["slds_K%d_d%d_T%d_B%d_run_inference" % (Ks, n, T, Bs)] = dict(
                    value=Bs / (ms * 1e-3), unit=UNIT, ms=ms,
                    roofline=_roof(8 * (3 * T * n) * Bs, ms, peak),       # node potentials in, one sample path out
                    cpu_baseline=cpu_rate("slds_port", (Ks, n, T), UNIT, secs=3.0, calib=1))
                del sJ, sh
                # GMM K=10 N=2 T=50: local mean field (whole fixed point in one launch), 256 minibatches per call
                from svae_b200.models import gmm as GM
                from svae_b200.distributions import dirichlet as DD, niw as NW
                Kg, Ng, Tg, Bg = 10, 2, 50, 256
                dg, ng = SY.gmm_global_natparam(Kg, Ng, np.random.default_rng(0))
                lg, gg = DD.expectedstats(tg(dg)), NW.expectedstats(tg(ng))
                gJ, gh = -0.5 * softplus(rnd(Bg, Tg, Ng)), 2.0 * rnd(Bg, Tg, Ng)
                li = torch.rand(Bg, Tg, Kg, dtype=torch.float64, device=dev, generator=gen)
                li = li / li.sum(-1, keepdim=True)
                ms = _time_gpu(torch, lambda: GM.meanfield_fixed_point_cuda(lg, gg, gJ, gh, li), iters=5)
                cfg["gmm_K%d_N%d_T%d_x%d_minibatches" % (Kg, Ng, Tg, Bg)] = dict(
                    value=Bg / (ms * 1e-3), unit="minibatches/s", ms=ms,
                    roofline=_roof(8 * (2 * Tg * Ng + 2 * Tg * Kg) * Bg, ms, peak),
                    cpu_baseline=cpu_rate("gmm_port", (Kg, Ng, Tg), "minibatches/s"))
                # small-batch / long-T regime: one sequence, sequential vs parallel in time
                from svae_b200.lds import lds_pit
                T3 = 20000
                J3, h3, e3 = -0.5 * softplus(rnd(1, T3, n)), rnd(1, T3, n), rnd(1, T3, 1, n)
                cfg["lds_d%d_T%d_B1_ms" % (n, T3)] = dict(
                    sequential=_time_gpu(torch, lambda: K.natural_lds_inference_general((gi, gp), (J3, h3), 1, e3), iters=3),
                    parallel_in_time=_time_gpu(torch, lambda: lds_pit.natural_lds_inference_parallel((gi, gp), (J3, h3), 1, e3), iters=3))
            except Exception as e:  # pragma: no cover
                cfg["error"] = repr(e)
            extra["configs"] = cfg
        if arm is not None:
            arm.close()

    if rank == 0:
        ab = alg_bytes_per_seq(n, T) * B + 8 * T * S * n * B     # + injected noise read
        achieved = ab / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("lds_estep_fwd_d10_T200_B4096")
            except Exception:
                traffic = None
        flops = T * (12 * n ** 3 + 10 * n ** 2) * B
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
            data="synthetic",
            config=dict(workload=WORKLOAD,
                        global_batch=world * B, parallelism="dp%d" % world,
                        l2="inputs per step (196 MB node potentials+noise, 1.4 GB stash) exceed the 126 MB L2; "
                           "%d input sets rotated" % NSETS,
                        step="fused filter+smoother+sampler kernel, 2-stage minibatch reduction of the global statistics"
                             + (", NCCL all-reduce of the %d-double accumulator" % (4 * n * n + n + 1) if world > 1 else "")),
            e2e=dict(value=world * B / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     ms_per_step=e2e_ms, api="svae_b200.models.lds.run_inference_host",
                     note="the model-level call make_gradfun makes (svae.py:21-22, models/lds.py:35-41) for a host-resident "
                          "minibatch: pinned node potentials in (%.0f MB), pinned samples + summed global statistics + local_kl out "
                          "(%.0f MB); the per-sequence node statistics are consumed by the local_kl contraction on the device (their "
                          "only consumer in run_inference) -- extra.e2e_all_outputs is the same pipeline with them copied back too; "
                          "pieces of %s sequences pipelined (one copy stream per direction, kernels on 3 streams), pinned buffers on "
                          "the GPU's NUMA node, persistent device buffers, noise drawn on device; floor = the host-to-device copy "
                          "(2.4 ms at the 55 GB/s measured) + the last piece's kernel (0.6 ms: 200 sequential steps) and copy"
                          % (h2d / 1e6, d2h / 1e6, RI_CHUNKS)),
            gpu_launches=int(launches),
            clocks=clk.summary(),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=traffic, peak_source=peak_src, kernel="lds_estep_packed_kernel<10,1,false>",
                          kernel_ms=kern_ms, algorithmic_bytes_per_launch=ab,
                          note="fp64 FMA / shared-memory-instruction / dependent-issue bound at ~30 flop/B, not HBM bound "
                               "(SURVEY.md 8d, DESIGN.md 3)",
                          # the second roofline SURVEY.md 8d asks for: algorithmic flops T(12n^3+10n^2) per sequence
                          # (the REFERENCE's operation count; this kernel's own formulation needs ~30% fewer) over the
                          # fp64 FMA peak 148 SMs x 64 FMA/clk x 2 x 1.965 GHz
                          fp64=dict(achieved_tflops=flops / (kern_ms * 1e-3) / 1e12,
                                    peak_tflops=148 * 64 * 2 * 1.965e9 / 1e12,
                                    frac=flops / (kern_ms * 1e-3) / (148 * 64 * 2 * 1.965e9))),
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
