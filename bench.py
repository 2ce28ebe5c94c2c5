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
                "sample": "%d training steps of %s at mini-batch %d through oracle/cpu_ref.cpp (C++ port of the reference's op semantics, "
                          "one 64-invocation work item per pool task, %d host threads)" % (steps, self.network, self.sample_batch, self.threads),
                "ms_per_step": ms_per_step}

    def close(self):
        self.program.close()
        self.env.close()


CPU_SAMPLE_BATCH = {"sentiment": 256, "conv-net": 1000, "conv-blur-net": 1000, "multi-hash": 16384, "siren": 16384, "relu": 16384, "relu-pe": 16384}


def cpu_baseline(network, optimizer, budget_s=15.0):
    """Bounded sample for the `cpu_baseline` object of the main arm: one warm-up step, then steps until ~budget_s."""
    ref = CpuReference(network, CPU_SAMPLE_BATCH.get(network, 1000), optimizer)
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
    ref = CpuReference(args.workload, CPU_SAMPLE_BATCH.get(args.workload, 1000), args.optimizer)
    for _ in range(args.warmup):
        ref.step()
    times = [ref.step() for _ in range(args.steps)]
    ms = float(np.mean(times)) * 1e3
    base = ref.describe(args.steps, ms)
    ref.close()
    print(json.dumps({
        "impl": "reference", "metric": "train samples/s", "value": base["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args), "sample_mini_batch": ref.sample_batch,
                                        "note": "multi-threaded C++ port of the reference's op semantics (oracle/cpu_ref.cpp) on the host cores; the "
                                                "reference's own Vulkan path cannot be built or run here (SURVEY.md section 0)"},
        "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_name(args):
    if args.workload == "sentiment":
        return "sentiment (vocabulary 4096, 32 words, embedding 128, LSTM 64)"
    return "fashion_mnist %s" % args.workload if args.workload in ("linear", "single-layer", "single-layer-dropout", "conv-net", "conv-blur-net") \
        else "image_fit %s" % args.workload


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
                    help="tf32: plain dense GEMMs on tcgen05 tensor cores (TF32 operands, FP32 accumulate); strict: all FP32 SIMT")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default="", help="write the per-kernel event timings here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import descent_b200 as d

    m = args.mini_batch or DEFAULT_BATCH.get(args.workload, 8192)
    env = d.Environment(local_rank)
    if world > 1:
        uid = [d.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        env.init_data_parallel(world, rank, uid[0])
    env.set_tf32(args.precision == "tf32")
    ex = env.example(args.workload, m, optimizer=args.optimizer)
    rng = np.random.default_rng(0x5EED5EED + 2)       # same initial weights on every rank
    params, _, _ = make_inputs(ex, rng, args.workload)
    batch_rng = np.random.default_rng(1000 + rank)    # a different shard of the global batch per rank
    from helpers import synthetic_batch
    x, y = synthetic_batch(ex, batch_rng)
    for pid, v in params.items():
        env.write(env.parameter(pid), v)
    env.write(ex.x, x)
    env.write(ex.y, y)
    x_pinned, y_pinned = d.pinned_array(x.size), d.pinned_array(y.size)
    x_pinned[:] = x.reshape(-1)
    y_pinned[:] = y.reshape(-1)

    def barrier():
        env.sync()
        if dist is not None:
            dist.barrier()

    import ctypes
    lib, ctx = d.lib, env.ctx()

    def timed(fn, steps):
        start, end = ctypes.c_void_p(), ctypes.c_void_p()
        lib.dsc_event_create(ctypes.byref(start))
        lib.dsc_event_create(ctypes.byref(end))
        barrier()
        lib.dsc_event_record(ctx, start)
        for s in range(steps):
            fn(s)
        lib.dsc_event_record(ctx, end)
        env.sync()
        ms = ctypes.c_float(0)
        lib.dsc_event_elapsed_ms(start, end, ctypes.byref(ms))
        barrier()
        total = ms.value
        if dist is not None:
            import torch
            t = torch.tensor([total], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total

    seeds = np.random.default_rng(5).integers(0, 2 ** 32, size=args.warmup + args.steps + 8)
    for s in range(args.warmup):
        env.run(ex.train_graph, int(seeds[s]))
    stats = env.graph_stats(ex.train_graph)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(lambda s: env.run(ex.train_graph, int(seeds[args.warmup + s])), args.steps)

    def upload():  # this step's share of host -> device traffic: one whole batch from pinned host memory
        env.prefetch_pinned(ex.x, x_pinned)
        env.prefetch_pinned(ex.y, y_pinned)

    def e2e_step(s):
        env.run(ex.train_graph, int(seeds[args.warmup + s]))  # consumes the batch uploaded during the previous step
        upload()                                               # batch s+1 crosses PCIe on the copy stream while step s computes
        env.read_parameter_scalar(ex.loss_sum)                 # device -> host read of step s's result (synchronises)
    upload()
    for s in range(2):
        e2e_step(s)
    e2e_ms = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    profile = env.profile(ex.train_graph, 1, 5)  # every rank: the step contains the gradient all-reduce
    barrier()
    if rank != 0:
        env.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    global_batch = m * world
    ms_per_step = total_ms / args.steps
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
                "algorithmic_bytes": top["bytes"], "peak_source": peak_source,
                "step": {"algorithmic_bytes": stats["algorithmic_bytes"], "flops": stats["flops"],
                         "achieved_gbs": stats["algorithmic_bytes"] / (ms_per_step * 1e-3) / 1e9,
                         "frac": stats["algorithmic_bytes"] / (ms_per_step * 1e-3) / 1e9 / hbm_peak}}
    if args.profile_json:
        with open(args.profile_json, "w") as f:
            json.dump({"workload": args.workload, "mini_batch": m, "ms_per_step_graph_replay": ms_per_step, "launches": profile}, f, indent=1)

    base = None
    if not args.no_cpu_baseline and world == 1:
        base = cpu_baseline(args.workload, args.optimizer)

    out = {
        "metric": "train samples/s", "value": global_batch / (ms_per_step * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 storage; tf32 tensor-core operands on dense GEMMs" if args.precision == "tf32" else "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "mini_batch_per_gpu": m, "global_batch": global_batch, "optimizer": args.optimizer,
                   "parallelism": "dp%d" % world,
                   "gemm_path": "tcgen05 tf32 operands, fp32 accumulate (TMA-fed dense GEMMs; halo-tiled implicit-GEMM conv forward / backward-input / weight-gradient); strict-fp32 streaming kernels for the 1-channel first conv and the tiny GEMMs" if args.precision == "tf32" else "strict-fp32 simt",
                   "cuda_graph": True,
                   "l2_policy": "working set per step (%.0f MB arena) exceeds the 126 MB L2" % (stats["arena_bytes"] / 1e6)
                   if stats["arena_bytes"] > 126e6 else "working set %.0f MB fits L2; no flush" % (stats["arena_bytes"] / 1e6)},
        "e2e": {"value": global_batch / (e2e_ms / args.steps * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(x.nbytes + y.nbytes),
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": stats["kernel_launches"] * args.steps,
        "kernels_per_step": stats["kernel_launches"],
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": base,
        "jit_ms": stats["jit_ms"],
    }
    print(json.dumps(out))
    env.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
