#!/usr/bin/env python
"""Measured parity margins on the GPU box: every golden fixture (recorded from the unmodified reference) replayed
step by step through conzic_gibbs_step in both arithmetic modes.  Writes a markdown table.

    python tools/parity_report.py > gpurun_out/parity.md
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_common as gc  # noqa: E402

FIX = ["seq_b2_n4_k8", "shuffle_b3_n5_k16_multi", "senti_shuffle_neg_b2_n4_k8", "peaked_seq_b2_n4_k32",
       "random_b2_n3_k8", "senti_seq_b2_n4_k8", "seq_b1_n10_k200"]
print("# Teacher-forced parity against the reference fixtures (tests/golden), measured on B200\n")
print("| fixture | mode | steps | max abs logit err | top-k id mismatches | max abs cosine err | max abs softmax_K err |"
      " winners checked | winner mismatches | min fused-score margin at a mismatch |")
print("|---|---|---|---|---|---|---|---|---|---|")
for prec in ("bf16x3", "bf16"):
    for name in FIX:
        g = gc.load_golden(name)
        case = g["case"]
        eng = gc.engine(prec, "tcgen05", case.get("peaked", False), case.get("multi", False))
        steps = [s for s in g["steps"] if not s.get("stale_logits")]
        g2 = dict(g, steps=steps)
        m = gc.replay_fixture(eng, g2)
        print(f"| {name} | {prec} | {m['steps']} | {m['logit_err']:.2e} | {m['topk_id_mismatch']} | {m['clip_ref_err']:.2e} | "
              f"{m['clip_score_err']:.2e} | {m['winner_checked']} | {m['winner_mismatch']} | {m['min_margin_at_mismatch']} |")
    gc.drop_engines()
