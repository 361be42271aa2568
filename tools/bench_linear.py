#!/usr/bin/env python
"""One linear-layer shape through conzic_debug_linear, GEMM-only device time from the library's CUDA events.
Used for A/B runs in one process and as the short command behind ncu captures of the persistent GEMM.

    python tools/bench_linear.py --M 75776 --N 1536 --K 512 --mode bf16 [--reps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=75776)
    ap.add_argument("--N", type=int, default=1536)
    ap.add_argument("--K", type=int, default=512)
    ap.add_argument("--mode", default="bf16", choices=["f32", "f32+resid", "bf16", "bf16+gelu"])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    a = ap.parse_args()
    import torch
    import gpu_common as gc
    eng = gc.engine("bf16", "tcgen05")
    A = torch.randn(a.M, a.K, device="cuda")
    W = torch.randn(a.N, a.K, device="cuda") * 0.05
    bias = torch.randn(a.N, device="cuda")
    resid = torch.randn(a.M, a.N, device="cuda") if a.mode == "f32+resid" else None
    act = {"f32": 0, "f32+resid": 0, "bf16": 16, "bf16+gelu": 17}[a.mode]
    # reference clock for this box: a plain library GEMM
    x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        x @ x
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        x @ x
    e1.record()
    torch.cuda.synchronize()
    ref_tf = 10 * 2 * 8192 ** 3 / (e0.elapsed_time(e1) / 1e3) / 1e12
    for _ in range(a.warm):
        eng.debug_linear(A, W, bias, resid, act)
    eng.profile(True)
    for _ in range(a.reps):
        eng.debug_linear(A, W, bias, resid, act)
    ms, work, n = eng.profile_read()["gemm"]
    eng.profile(False)
    print(json.dumps(dict(M=a.M, N=a.N, K=a.K, mode=a.mode, ms=round(ms / n, 4),
                          tflops=round(work / (ms / 1e3) / 1e12, 1), cublas_8192_tflops=round(ref_tf, 1))))


if __name__ == "__main__":
    main()
