#!/usr/bin/env python
"""Fused MLP kernel (conzic_debug_mlp) device time from the library's CUDA events."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import gpu_common as gc  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
eng = gc.engine("bf16", "tcgen05")
H, F = 512, 2048
X = torch.randn(M, H, device="cuda")
W1 = torch.randn(F, H, device="cuda") * 0.04
W2 = torch.randn(H, F, device="cuda") * 0.02
b1 = torch.randn(F, device="cuda") * 0.1
b2 = torch.randn(H, device="cuda") * 0.1
for _ in range(3):
    eng.debug_mlp(X, W1, b1, W2, b2, 1)
eng.profile(True)
for _ in range(reps):
    eng.debug_mlp(X, W1, b1, W2, b2, 1)
ms, work, n = eng.profile_read()["gemm"]
eng.profile(False)
print(json.dumps(dict(kernel="mlp_persist", M=M, ms=round(ms / n, 4), tflops=round(work / (ms / 1e3) / 1e12, 1))))
