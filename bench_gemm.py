#!/usr/bin/env python
"""GEMM % of tensor peak: dsc_gemm_tf32 on a synthetic M=N=K GEMM, next to cuBLAS TF32 (torch.matmul) on the same
device as the measured denominator (TF32 peak is not in MEASURED_PEAKS.json)."""
import argparse
import ctypes
import json
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--layouts", default="10,11,00,01")
    ap.add_argument("--no-cublas", action="store_true")
    args = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()
    import descent_b200 as d
    env = d.Environment(0)
    lib, ctx = d.lib, env.ctx()
    n = args.size
    rng = np.random.default_rng(0)
    host = rng.standard_normal((n, n)).astype(np.float32)
    bufs = []
    for _ in range(3):
        h = ctypes.c_uint64(0)
        assert lib.dsc_alloc(ctx, ctypes.c_size_t(host.nbytes), ctypes.byref(h)) == 0
        assert lib.dsc_upload(ctx, h, ctypes.c_size_t(0), host.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(host.nbytes), ctypes.c_size_t(0), 0) == 0
        bufs.append(h)
    results = {}
    for lay in args.layouts.split(","):
        a_is_mk, b_is_kn = int(lay[0]), int(lay[1])
        for _ in range(3):
            assert lib.dsc_gemm_tf32(ctx, bufs[0], bufs[1], bufs[2], ctypes.c_int64(n), ctypes.c_int64(n), ctypes.c_int64(n), a_is_mk, b_is_kn) == 0, lib.dsc_last_error()
        start, end = ctypes.c_void_p(), ctypes.c_void_p()
        lib.dsc_event_create(ctypes.byref(start)); lib.dsc_event_create(ctypes.byref(end))
        env.sync()
        best = 1e9
        total = 0.0
        for _ in range(args.iters):
            lib.dsc_event_record(ctx, start)
            lib.dsc_gemm_tf32(ctx, bufs[0], bufs[1], bufs[2], ctypes.c_int64(n), ctypes.c_int64(n), ctypes.c_int64(n), a_is_mk, b_is_kn)
            lib.dsc_event_record(ctx, end)
            ms = ctypes.c_float(0)
            lib.dsc_event_elapsed_ms(start, end, ctypes.byref(ms))
            best = min(best, ms.value)
            total += ms.value
        results["a_mk=%d,b_kn=%d" % (a_is_mk, b_is_kn)] = {"best_ms": best, "mean_ms": total / args.iters, "best_tflops": 2.0 * n ** 3 / (best * 1e-3) / 1e12,
                                                      "mean_tflops": 2.0 * n ** 3 / (total / args.iters * 1e-3) / 1e12}
    cublas = None
    if not args.no_cublas:
        import torch
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(args.iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); a @ b; e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        cublas = {"best_ms": best, "best_tflops": 2.0 * n ** 3 / (best * 1e-3) / 1e12}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    out = {"metric": "GEMM % tensor peak", "size": n, "dtype": "tf32 operands, f32 accumulate", "descent_b200": results, "cublas_tf32": cublas,
           "derived_tf32_peak_tflops": peaks.get("bf16_tflops", 1590.0) / 2.0, "peak_note": "half of the measured bf16 burst GEMM rate (MEASURED_PEAKS.json); cuBLAS TF32 measured alongside"}
    best_ours = max(r["best_tflops"] for r in results.values())
    out["frac_of_derived_peak"] = best_ours / out["derived_tf32_peak_tflops"]
    if cublas:
        out["frac_of_cublas_tf32"] = best_ours / cublas["best_tflops"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
