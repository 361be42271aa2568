#!/usr/bin/env python
"""BASELINE config 5: candidate_k in {50,200,512} x sentence_len in {5,10,25} on one GPU (or one rank of many):
captions/s, algorithmic TFLOP per caption, achieved fraction of the GEMM roofline.

    python tools/sweep.py [--batch 64] [--sweeps 5] [--reps 2] [--out gpurun_out/sweep.jsonl]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from conzic_b200 import synth  # noqa: E402
from conzic_b200.engine import Engine  # noqa: E402


def algo_tflop_per_caption(n, K, sweeps):
    """every candidate caption encoded in full by the CLIP text tower."""
    L = n + 5
    clip_tok = K * (n * (n + 11) / 2 + (sweeps - 1) * n * (n + 5))      # tokens over all steps (first sweep grows)
    clip = clip_tok * 75_497_472 + sweeps * n * K * 524_288
    # causal attention: 12288 * T per token
    att = K * 12_288 * (sum((ii + 6) ** 2 for ii in range(n)) + (sweeps - 1) * n * (n + 5) ** 2)
    bert = sweeps * n * (L * (169_869_312 + 36_864 * L) + 48_061_440)
    return (clip + att + bert) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--sweeps", type=int, default=5)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--ks", default="50,200,512")
    ap.add_argument("--lens", default="5,10,25")
    a = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}
    eng = Engine(synth.make_bert_state_dict(0), synth.make_clip_state_dict(0, vision=False), precision="bf16")
    eng.set_bert2clip(*synth.build_bert2clip_table(False))
    B = a.batch
    img = torch.nn.functional.normalize(torch.randn(B, 512, device="cuda", generator=torch.Generator("cuda").manual_seed(1)), dim=-1)
    tok = synth.SynthBertTokenizer()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    for n in [int(x) for x in a.lens.split(",")]:
        for K in [int(x) for x in a.ks.split(",")]:
            init = torch.tensor([tok.encode(synth.SYNTH_PROMPT + "[MASK]" * n)] * B, device="cuda")
            clip_ref = torch.zeros(B, device="cuda")

            def run():
                inp = init.clone()
                tm = synth.make_token_mask("cuda")
                holds = [False] + [True] * 3 + [False] * (n + 1)
                for _ in range(a.sweeps):
                    for ii in range(n):
                        pos = 4 + ii
                        eng.gibbs_step(inp, tm, img, pos, ii == n - 1, K, 0.1, 0.02, 2.0, sum(holds[:pos]),
                                       sum(holds[pos + 1:]), out_clip_ref=clip_ref)
                        holds[pos] = True
                return inp

            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            eng.profile(True)
            out = run()
            pr = eng.profile_read()
            eng.profile(False)
            g_ms, g_fl, _ = pr["gemm"]
            alg = algo_tflop_per_caption(n, K, a.sweeps)
            cps = B / (ms / 1e3)
            rec = dict(sentence_len=n, candidate_k=K, batch=B, sweeps=a.sweeps, ms_per_batch=round(ms, 2),
                       captions_per_s=round(cps, 2), algorithmic_tflop_per_caption=round(alg, 3),
                       algorithmic_tflops=round(cps * alg, 1),
                       executed_gemm_tflops=round(g_fl / (g_ms / 1e3) / 1e12, 1),
                       executed_gemm_frac_of_peak=round(g_fl / (g_ms / 1e3) / 1e12 / peaks["bf16_tflops_sustained"], 3),
                       gemm_share_of_step=round(g_ms / sum(v[0] for v in pr.values()), 3),
                       breakdown_ms={k: round(v[0], 2) for k, v in pr.items()},
                       legal_winners=bool((out[:, 4:4 + n] >= 1012).all()))
            print(json.dumps(rec), flush=True)
            with open(a.out, "a") as fh:
                fh.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
