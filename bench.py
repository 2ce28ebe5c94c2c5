#!/usr/bin/env python
"""Headline benchmark: training samples/s of the reference's example networks on the CUDA backend.

    python bench.py --gpus N --steps K --warmup W [--workload conv-net] [--mini-batch M]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (N > 1)
    python bench.py --impl reference ...     (the CPU port of the reference path, see oracle/)

One JSON line on rank 0.  `value` = global samples / max-over-ranks device time for K replayed steps with the
batch resident in HBM; `e2e` = the same through the public API with host buffers: every step uploads one batch
from pinned host memory (Environment.prefetch_pinned, overlapped with the previous step on a copy stream), runs,
and reads the loss back.  `roofline` describes the kernel with the largest share of the step (CUDA events per
launch on the context's stream), against MEASURED_PEAKS.json.  `cpu_baseline` times the multi-threaded C++ port of the
reference's op semantics (oracle/cpu_ref.cpp) on a bounded sample of the same workload on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md
DEFAULT_BATCH = {"conv-net": 8192, "conv-blur-net": 8192, "linear": 8192, "single-layer": 8192, "single-layer-dropout": 8192,
                 "multi-hash": 262144, "siren": 65536, "relu": 65536, "relu-pe": 65536, "sentiment": 256}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("hbm_gbs", FALLBACK_HBM_GBS), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(ex, rng, network):
    from helpers import init_example_params, synthetic_batch
    params = init_example_params(ex, rng, siren=(network == "siren"))
    x, y = synthetic_batch(ex, rng)
    return params, x, y


class CpuReference:
    """The multi-threaded C++ port of the reference's op semantics (oracle/cpu_ref.cpp) on a bounded sample of the
    workload: host-only graph build, then whole training steps of the exported graph (after the reference's graph
    passes, so views are folded) on every host thread the box has."""

    def __init__(self, network, sample_batch, optimizer):
        import descent_b200 as d
        from oracle import cpu_ref
        self.network, self.sample_batch = network, sample_batch
        self.env = d.Environment(-1)
        ex = self.env.example(network, sample_batch, optimizer=optimizer)
        rng = np.random.default_rng(0x5EED5EED)
        self.params, x, y = make_inputs(ex, rng, network)
        graph = ex.train_graph.export_json()
        for node in graph["nodes"]:
            if node["op"] == "Input" and node["parameter"] not in self.params:
                shape = self.env.parameter(node["parameter"]).shape()
                self.params[node["parameter"]] = np.full(shape, 1.0 / 16.0, np.float32)
        self.params[ex.x.id], self.params[ex.y.id] = x, y
        self.threads = cpu_ref.hardware_threads()
        self.program = cpu_ref.Program(graph)
        self.seed = 0

    def step(self):
        """One training step; returns its wall time in seconds (measured inside the library, around the graph run)."""
        out, seconds = self.program.run(self.params, self.seed, self.threads)
        self.params.update({pid: v.reshape(-1) for pid, v in out.items()})
        self.seed += 1
        return seconds

    def describe(self, steps, ms_per_step):
        return {"value": self.sample_batch / (ms_per_step * 1e-3), "unit": "samples/s", "cores": self.threads, "kind": "port",
                "sample": "%d training steps of %s on a %d-sample slice of the mini-batch through oracle/cpu_ref.cpp (C++ port of the reference's op semantics, "
                          "one 64-invocation work item per pool task, %d host threads)" % (steps, self.network, self.sample_batch, self.threads),
                "ms_per_step": ms_per_step}

    def close(self):
        self.program.close()
        self.env.close()


# The CPU arm is configured like the GPU arm (same workload, per-GPU mini-batch, optimiser); each of its steps processes a
# bounded SAMPLE of that mini-batch -- this many samples -- so that K + W steps end within a few minutes (a whole conv-net
# mini-batch of 8192 takes ~11 s per step on 16 host threads).  samples/s is a per-sample rate, so the two arms compare.
CPU_SAMPLE_BATCH = {"sentiment": 256, "conv-net": 1024, "conv-blur-net": 1024, "multi-hash": 16384, "siren": 16384, "relu": 16384, "relu-pe": 16384}


def cpu_baseline(network, optimizer, budget_s=15.0):
    """Bounded sample for the `cpu_baseline` object of the main arm: one warm-up step, then steps until ~budget_s."""
    ref = CpuReference(network, CPU_SAMPLE_BATCH.get(network, 1024), optimizer)
    ref.step()
    times = []
    while sum(times) < budget_s and len(times) < 20:
        times.append(ref.step())
    base = ref.describe(len(times), float(np.mean(times)) * 1e3)
    ref.close()
    return base


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    m = args.mini_batch or DEFAULT_BATCH.get(args.workload, 8192)
    ref = CpuReference(args.workload, min(m, CPU_SAMPLE_BATCH.get(args.workload, 1024)), args.optimizer)
    for _ in range(args.warmup):
        ref.step()
    times = [ref.step() for _ in range(args.steps)]
    ms = float(np.mean(times)) * 1e3
    base = ref.describe(args.steps, ms)
    ref.close()
    print(json.dumps({
        "impl": "reference", "metric": "train samples/s", "value": base["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "mini_batch_per_gpu": m, "global_batch": m * world, "optimizer": args.optimizer,
                   "parallelism": "dp%d" % world,
                   "note": "multi-threaded C++ port of the reference's op semantics (oracle/cpu_ref.cpp) on rank 0's host cores; each step processes a %d-sample "
                           "slice of the mini-batch (bounded sample); the reference's own Vulkan path cannot be built or run here (SURVEY.md section 0)" % ref.sample_batch},
        "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_name(args):
    if args.workload == "sentiment":
        return "sentiment (vocabulary 4096, 32 words, embedding 128, LSTM 64)"
    return "fashion_mnist %s" % args.workload if args.workload in ("linear", "single-layer", "single-layer-dropout", "conv-net", "conv-blur-net") \
        else "image_fit %s" % args.workload


IMAGE_WIDTH = {"multi-hash": 1024, "siren": 512, "relu": 512, "relu-pe": 512}  # BASELINE.json configs 4 and 5
L2_BYTES = 126e6


class Rig:
    """Process-wide state shared by every workload measured in one bench.py run."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None

    def barrier(self, env):
        env.sync()
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, value):
        if self.dist is None:
            return value
        import torch
        t = torch.tensor([value], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure_workload(d, rig, workload, m, precision, steps, warmup, optimizer, want_profile):
    """One workload on this rank's GPU (data-parallel over all ranks when world > 1): K graph replays with the batch
    resident (`value`), K steps through the public API with a host batch uploaded and the loss read back every step
    (`e2e`), and per-kernel event timings of an eager pass (`profile`)."""
    import ctypes
    from helpers import synthetic_batch
    lib = d.lib
    env = d.Environment(rig.local_rank)
    ctx = env.ctx()
    if rig.world > 1:
        uid = [d.nccl_unique_id() if rig.rank == 0 else None]
        rig.dist.broadcast_object_list(uid, src=0)
        env.init_data_parallel(rig.world, rig.rank, uid[0])
    env.set_tf32(precision == "tf32")
    ex = env.example(workload, m, optimizer=optimizer)
    rng = np.random.default_rng(0x5EED5EED + 2)           # same initial weights on every rank
    params, _, _ = make_inputs(ex, rng, workload)
    batch_rng = np.random.default_rng(1000 + rig.rank)    # a different shard of the global batch per rank
    x, y = synthetic_batch(ex, batch_rng, IMAGE_WIDTH.get(workload, 512))
    for pid, v in params.items():
        env.write(env.parameter(pid), v)
    env.write(ex.x, x)
    env.write(ex.y, y)
    x_pinned, y_pinned = d.pinned_array(x.size), d.pinned_array(y.size)
    x_pinned[:] = x.reshape(-1)
    y_pinned[:] = y.reshape(-1)
    seeds = np.random.default_rng(5).integers(0, 2 ** 32, size=warmup + steps + 8)
    for s in range(warmup):
        env.run(ex.train_graph, int(seeds[s]))
    stats = env.graph_stats(ex.train_graph)
    # L2 policy: a step whose working set (the arena of intermediates) exceeds the 126 MB L2 evicts its own data; a smaller
    # one is timed step by step with a 256 MB fill between the timed steps (outside the event pairs)
    flush = stats["arena_bytes"] < 1.5 * L2_BYTES
    flush_buf = ctypes.c_uint64(0)
    if flush:
        assert lib.dsc_alloc(ctx, ctypes.c_size_t(256 << 20), ctypes.byref(flush_buf)) == 0

    def timed(fn, count):
        start, end = ctypes.c_void_p(), ctypes.c_void_p()
        lib.dsc_event_create(ctypes.byref(start))
        lib.dsc_event_create(ctypes.byref(end))
        ms = ctypes.c_float(0)
        rig.barrier(env)
        if not flush:
            lib.dsc_event_record(ctx, start)
            for s in range(count):
                fn(s)
            lib.dsc_event_record(ctx, end)
            env.sync()
            lib.dsc_event_elapsed_ms(start, end, ctypes.byref(ms))
            total = ms.value
        else:
            total = 0.0
            for s in range(count):
                lib.dsc_fill_u32(ctx, flush_buf, ctypes.c_size_t(0), ctypes.c_uint32(s), ctypes.c_size_t((256 << 20) // 4))
                lib.dsc_event_record(ctx, start)
                fn(s)
                lib.dsc_event_record(ctx, end)
                env.sync()
                lib.dsc_event_elapsed_ms(start, end, ctypes.byref(ms))
                total += ms.value
        rig.barrier(env)
        return rig.max_over_ranks(total)

    total_ms = timed(lambda s: env.run(ex.train_graph, int(seeds[warmup + s])), steps)

    def upload():  # this step's share of host -> device traffic: one whole batch from pinned host memory
        env.prefetch_pinned(ex.x, x_pinned)
        env.prefetch_pinned(ex.y, y_pinned)

    def e2e_step(s):
        env.run(ex.train_graph, int(seeds[warmup + s]))  # consumes the batch uploaded during the previous step
        upload()                                          # batch s+1 crosses PCIe on the copy stream while step s computes
        env.read_parameter_scalar(ex.loss_sum)            # device -> host read of step s's result (synchronises)
    upload()
    for s in range(2):
        e2e_step(s)
    e2e_ms = timed(e2e_step, steps)
    profile = env.profile(ex.train_graph, 1, 5) if want_profile else None  # every rank: the step contains the gradient all-reduce
    rig.barrier(env)
    loss = env.read_parameter_scalar(ex.loss_sum)
    if flush:
        lib.dsc_free(ctx, flush_buf)
    env.close()
    hbm_peak, _ = peaks()
    ms_per_step = total_ms / steps
    global_batch = m * rig.world
    return {
        "value": global_batch / (ms_per_step * 1e-3), "unit": "samples/s", "ms_per_step": ms_per_step, "mini_batch_per_gpu": m,
        "global_batch": global_batch, "kernels_per_step": stats["kernel_launches"], "precision": precision,
        "e2e": {"value": global_batch / (e2e_ms / steps * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(x.nbytes + y.nbytes),
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / steps},
        "step_roofline": {"algorithmic_bytes": stats["algorithmic_bytes"], "flops": stats["flops"],
                          "achieved_gbs": stats["algorithmic_bytes"] / (ms_per_step * 1e-3) / 1e9,
                          "frac": stats["algorithmic_bytes"] / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                          # a fused launch (the dense-chain MLP kernel) removes traffic instead of moving it faster: the same
                          # fraction with that launch counted as the bytes of the kernels it replaces
                          "unfused_algorithmic_bytes": stats.get("unfused_algorithmic_bytes", stats["algorithmic_bytes"]),
                          "frac_of_unfused_traffic": stats.get("unfused_algorithmic_bytes", stats["algorithmic_bytes"]) / (ms_per_step * 1e-3) / 1e9 / hbm_peak},
        "l2_policy": ("256 MB fill between timed steps (working set %.0f MB is L2-sized)" if flush else
                      "working set per step (%.0f MB arena) exceeds the 126 MB L2") % (stats["arena_bytes"] / 1e6),
        "loss_sum_finite": bool(np.isfinite(loss)), "jit_ms": stats["jit_ms"], "_profile": profile, "_stats": stats,
    }


def gemm_section(d, rig):
    """GEMM % of tensor peak (BASELINE.json metric), through the public graph path: c = a.matmul(b) lowers to
    MatMul[r = ceil(K / 1024)] + Reduce, which the backend absorbs into one dense tcgen05 TF32 GEMM (TMA-fed, TMEM
    accumulators).  Best of 10 and 30 back to back, beside cuBLAS TF32 (torch.matmul) on the same device, with the clocks
    seen and a correctness check of sampled entries against the float64 product of TF32-truncated operands."""
    import ctypes
    lib = d.lib
    out = {}
    sampler = ClockSampler(rig.local_rank)
    sampler.start()
    for name, (M, K, N) in (("8192^3", (8192, 8192, 8192)), ("conv-shaped [2^20,1152]x[1152,128]", (1 << 20, 1152, 128))):
        env = d.Environment(rig.local_rank)
        env.set_tf32(True)
        ctx = env.ctx()
        a, b, c = env.static_parameter([M, K], "a"), env.static_parameter([K, N], "b"), env.static_parameter([M, N], "c")
        scope = env.scope()
        scope.write_parameter_value(c, scope.parameter_value(a).matmul(scope.parameter_value(b)))
        graph = scope.build_graph()
        rng = np.random.default_rng(M + N + K)
        ha, hb = rng.standard_normal((M, K), dtype=np.float32), rng.standard_normal((K, N), dtype=np.float32)
        env.write(a, ha)
        env.write(b, hb)
        for _ in range(3):
            env.run(graph, 0)
        labels = [t["label"] for t in env.profile(graph, 0, 1)]
        start, end = ctypes.c_void_p(), ctypes.c_void_p()
        lib.dsc_event_create(ctypes.byref(start))
        lib.dsc_event_create(ctypes.byref(end))
        ms = ctypes.c_float(0)
        best = 1e9
        for _ in range(10):
            env.sync()
            lib.dsc_event_record(ctx, start)
            env.run(graph, 0)
            lib.dsc_event_record(ctx, end)
            env.sync()
            lib.dsc_event_elapsed_ms(start, end, ctypes.byref(ms))
            best = min(best, ms.value)
        lib.dsc_event_record(ctx, start)
        for _ in range(30):
            env.run(graph, 0)
        lib.dsc_event_record(ctx, end)
        env.sync()
        lib.dsc_event_elapsed_ms(start, end, ctypes.byref(ms))
        sustained = ms.value / 30
        got = env.read(c)
        from oracle.interp import tf32_operand  # checker only
        idx = rng.integers(0, [M, N], size=(256, 2))
        ta, tb = tf32_operand(ha[idx[:, 0]], "trunc").astype(np.float64), tf32_operand(np.ascontiguousarray(hb[:, idx[:, 1]].T), "trunc").astype(np.float64)
        exact = (ta * tb).sum(1)
        err = float(np.abs(got[idx[:, 0], idx[:, 1]] - exact).max() / np.abs(exact).max())
        flops = 2.0 * M * N * K
        out[name] = {"m": M, "n": N, "k": K, "best_ms": best, "tflops": flops / (best * 1e-3) / 1e12, "sustained_ms": sustained,
                     "sustained_tflops": flops / (sustained * 1e-3) / 1e12, "max_rel_err_vs_tf32_exact": err, "kernels": labels}
        env.close()
    clocks = sampler.stop()
    cublas = None
    try:
        import torch
        torch.backends.cuda.matmul.allow_tf32 = True
        ta, tb = torch.randn(8192, 8192, device="cuda"), torch.randn(8192, 8192, device="cuda")
        for _ in range(3):
            ta @ tb
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ta @ tb
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        cublas = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        del ta, tb
    except Exception as exc:  # torch without CUDA: leave the library denominator out
        cublas = None
        out["cublas_error"] = str(exc)[:200]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16 = json.load(open(peaks_path)).get("bf16_tflops") if os.path.exists(peaks_path) else None
    big = out["8192^3"]
    return {"tflops": big["tflops"], "sustained_tflops": big["sustained_tflops"], "cublas_tf32_tflops": cublas,
            "frac_of_measured_tf32_peak": big["tflops"] / cublas if cublas else None,
            "frac_of_half_measured_bf16_peak": big["tflops"] / (bf16 / 2.0) if bf16 else None,
            "tf32_peak_note": "measured denominator = cuBLAS TF32 8192^3 (torch.matmul, allow_tf32) best of 10 in this run; MEASURED_PEAKS.json has no TF32 entry, "
                              "half of its bf16 burst rate is the derived figure",
            "max_rel_err_vs_tf32_exact": big["max_rel_err_vs_tf32_exact"], "clocks": clocks, "shapes": out,
            "path": "Array.matmul -> MatMul[r=8]+Reduce absorbed -> dsc_gemm_tf32 (TMA + tcgen05.mma kind::tf32 + TMEM)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="descent_b200", choices=["descent_b200", "reference"])
    ap.add_argument("--workload", default="conv-net")
    ap.add_argument("--mini-batch", type=int, default=0, help="per-GPU mini-batch (0 = workload default)")
    ap.add_argument("--optimizer", default="adam")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "strict"],
                    help="tf32: GEMMs / convolutions on tcgen05 tensor cores (TF32 operands, FP32 accumulate); strict: all FP32 SIMT")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE.json workloads, the strict-FP32 run and the GEMM section")
    ap.add_argument("--profile-json", default="", help="write the per-kernel event timings here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rig = Rig(args)
    rank, world = rig.rank, rig.world
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rig.local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", rig.local_rank))
        dist.barrier()
        rig.dist = dist
    import descent_b200 as d

    m = args.mini_batch or DEFAULT_BATCH.get(args.workload, 8192)
    sampler = ClockSampler(rig.local_rank)
    if rank == 0:
        sampler.start()
    main_res = measure_workload(d, rig, args.workload, m, args.precision, args.steps, args.warmup, args.optimizer, want_profile=True)
    clocks = sampler.stop() if rank == 0 else None

    # the other metrics BASELINE.json names, in the same process at the same N: hash-MLP / SIREN / ReLU+PE samples/s, the
    # strict-FP32 (1e-5 path) number of the headline workload, and GEMM % of tensor peak (N = 1 only)
    extras, strict, gemm = {}, None, None
    if not args.no_extras and args.workload == "conv-net":
        for name in ("multi-hash", "siren", "relu-pe"):
            r = measure_workload(d, rig, name, DEFAULT_BATCH[name], args.precision, args.steps, 3, "adam", want_profile=False)
            r["workload"] = workload_name(argparse.Namespace(workload=name)) + (" (random pixels of a synthetic %dx%d RGB image)" % (IMAGE_WIDTH[name], IMAGE_WIDTH[name]))
            extras[name] = {k: v for k, v in r.items() if not k.startswith("_")}
        r = measure_workload(d, rig, args.workload, m, "strict", max(5, args.steps // 2), 3, args.optimizer, want_profile=False)
        strict = {k: v for k, v in r.items() if not k.startswith("_")}
        if world == 1:
            gemm = gemm_section(d, rig)
    if rank != 0:
        if rig.dist is not None:
            rig.dist.destroy_process_group()
        return

    profile, stats = main_res["_profile"], main_res["_stats"]
    ms_per_step = main_res["ms_per_step"]
    hbm_peak, peak_source = peaks()
    by_entry = {}
    for t in profile:
        if not t["entry"]:
            continue
        e = by_entry.setdefault(t["entry"], {"label": t["label"], "ms": 0.0, "bytes": t["bytes"], "flops": t["flops"], "launches": 0})
        e["ms"] += t["ms"]
        e["launches"] += 1
    profiled_total = sum(t["ms"] for t in profile)
    top_entry, top = max(by_entry.items(), key=lambda kv: kv[1]["ms"])
    achieved = top["bytes"] / (top["ms"] / top["launches"] * 1e-3) / 1e9
    traffic = None
    traffic_path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_path):
        with open(traffic_path) as f:
            traffic = json.load(f).get("%s/%d/%s" % (args.workload, m, top["label"]))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "kernel": top["label"], "entry": top_entry, "kernel_ms": top["ms"] / top["launches"], "share_of_step": top["ms"] / profiled_total,
                "algorithmic_bytes": top["bytes"], "peak_source": peak_source, "step": main_res["step_roofline"]}
    if args.profile_json:
        with open(args.profile_json, "w") as f:
            json.dump({"workload": args.workload, "mini_batch": m, "ms_per_step_graph_replay": ms_per_step, "launches": profile}, f, indent=1)

    base = None
    if not args.no_cpu_baseline and world == 1:
        base = cpu_baseline(args.workload, args.optimizer)

    out = {
        "metric": "train samples/s", "value": main_res["value"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 storage; tf32 tensor-core operands on GEMMs / convolutions" if args.precision == "tf32" else "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "mini_batch_per_gpu": m, "global_batch": main_res["global_batch"], "optimizer": args.optimizer,
                   "parallelism": "dp%d" % world,
                   "gemm_path": "tcgen05 tf32 operands, fp32 accumulate (TMA-fed dense GEMMs; halo-tiled implicit-GEMM conv forward / backward-input / weight-gradient); strict-fp32 streaming kernels for the 1-channel first conv and the tiny GEMMs" if args.precision == "tf32" else "strict-fp32 simt",
                   "cuda_graph": True, "l2_policy": main_res["l2_policy"]},
        "e2e": main_res["e2e"],
        "gpu_launches": main_res["kernels_per_step"] * args.steps,
        "kernels_per_step": main_res["kernels_per_step"],
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": base,
        "jit_ms": main_res["jit_ms"],
    }
    if extras:
        out["workloads"] = extras
    if strict:
        out["strict"] = strict
    if gemm:
        out["gemm"] = gemm
    print(json.dumps(out))
    if rig.dist is not None:
        rig.dist.destroy_process_group()


if __name__ == "__main__":
    main()
