#!/usr/bin/env python
"""Short, fixed workload for ncu captures and CUDA-event breakdowns: Gibbs steps of BASELINE config 2
(B=64, K=200, len=10) at a chosen caption position of a later sweep (all other positions hold a word).

    python tools/profile_step.py [--ii 3] [--steps 1] [--warm 1] [--precision bf16] [--breakdown]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthetic import synth  # noqa: E402
from conzic_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ii", type=int, default=3)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--topk", type=int, default=200)
    ap.add_argument("--len", type=int, default=10)
    ap.add_argument("--precision", default="certified")
    ap.add_argument("--breakdown", action="store_true")
    a = ap.parse_args()
    B, n, K = a.batch, a.len, a.topk
    eng = Engine(synth.make_bert_state_dict(0), synth.make_clip_state_dict(0, vision=False), precision=a.precision)
    eng.set_bert2clip(*synth.build_bert2clip_table(False))
    img = torch.nn.functional.normalize(torch.randn(B, 512, device="cuda"), dim=-1)
    base = torch.tensor([[101, 3746, 1997, 1037] + [2000 + 7 * j for j in range(n)] + [102]] * B, device="cuda")
    tm = synth.make_token_mask("cuda")

    def step():
        inp = base.clone()
        eng.gibbs_step(inp, tm, img, 4 + a.ii, a.ii == n - 1, K, 0.1, 0.02, 2.0, 3 + a.ii, n - 1 - a.ii)

    for _ in range(a.warm):
        step()
    torch.cuda.synchronize()
    if a.breakdown:
        eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    out = {"B": B, "K": K, "len": n, "ii": a.ii, "precision": a.precision, "steps": a.steps,
           "ms_per_step": e0.elapsed_time(e1) / a.steps, "launches_per_step": (eng.launch_count() - l0) / a.steps}
    if a.breakdown:
        pr = eng.profile_read()
        out["phases_ms_per_step"] = {k: (round(v[0] / a.steps, 4), v[1] // a.steps) for k, v in eng.profile_read_phases().items()}
        out["breakdown_ms_per_step"] = {k: round(v[0] / a.steps, 4) for k, v in pr.items()}
        out["launches"] = {k: v[2] // a.steps for k, v in pr.items()}
        g = pr["gemm"]
        out["gemm_tflops"] = g[1] / (g[0] / 1e3) / 1e12 if g[0] else None
        eng.profile(False)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
