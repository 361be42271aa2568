#!/usr/bin/env python
"""Same-box comparison of engine variants selected by environment switches (read when the context is created):
full-size Gibbs steps of BASELINE config 2 (B=64, K=200, len=10), CUDA-event time per step at several caption
positions plus agreement of cosines / winners with the first (default) variant.

    python tools/exp_variants.py "" "CONZIC_WIDE_LN=1" "CONZIC_WIDE_LN=2"
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from conzic_b200 import synth  # noqa: E402
from conzic_b200.engine import Engine  # noqa: E402

B, n, K = 64, 10, 200
POSITIONS = tuple(int(x) for x in os.environ.get("EXP_POSITIONS", "0,5,9").split(","))


def run(envspec, bert_sd, clip_sd, table, base):
    keys = []
    for kv in envspec.split():
        k, v = kv.split("=")
        os.environ[k] = v
        keys.append(k)
    eng = Engine(bert_sd, clip_sd, precision="bf16")
    eng.set_bert2clip(*table)
    img = torch.nn.functional.normalize(torch.randn(B, 512, generator=torch.Generator().manual_seed(9)), dim=-1).cuda()
    ids = torch.tensor([[101, 3746, 1997, 1037] + [2000 + 7 * j for j in range(n)] + [102]] * B, device="cuda")
    tm = synth.make_token_mask("cuda")
    out = {"variant": envspec or "default", "ms": {}, "agree": {}}
    for ii in POSITIONS:
        def step(trace=False):
            inp = ids.clone()
            r = eng.gibbs_step(inp, tm, img, 4 + ii, ii == n - 1, K, 0.1, 0.02, 2.0, 3 + ii, n - 1 - ii, trace=trace)
            return inp, r
        inp, (cr, _, tr) = step(True)
        torch.cuda.synchronize()
        key = f"ii{ii}"
        got = (inp[:, 4 + ii].cpu(), tr["clip_ref"].cpu())
        if base is not None:
            w0, c0 = base[key]
            out["agree"][key] = {"winners_equal": int((w0 == got[0]).sum()), "of": B,
                                 "max_abs_dcos": float((c0 - got[1]).abs().max())}
        out.setdefault("_raw", {})[key] = got
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        best = 1e9
        for _rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                step()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 4)
        out["ms"][key] = round(best, 3)
        if ii == POSITIONS[0]:  # per-category CUDA-event sums of one step (adds event overhead; shares only)
            eng.profile(True)
            step()
            torch.cuda.synchronize()
            out["breakdown_ms"] = {k: round(v[0], 3) for k, v in eng.profile_read().items() if v[0] > 0.05}
            eng.profile(False)
    eng.close()
    for k in keys:
        os.environ.pop(k)
    torch.cuda.empty_cache()
    return out


def main():
    specs = sys.argv[1:] or [""]
    bert_sd, clip_sd = synth.make_bert_state_dict(0), synth.make_clip_state_dict(0, vision=False)
    table = synth.build_bert2clip_table(False)
    base = None
    for rep in range(2):  # two rounds so drift on the box shows
        for spec in specs:
            r = run(spec, bert_sd, clip_sd, table, base)
            raw = r.pop("_raw")
            if base is None:
                base = raw
            r["round"] = rep
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
