#!/usr/bin/env python
"""captions/s of generate_caption (BASELINE config 2 sizes, host pixels in, strings out) for the three candidate-text
paths: the plain synthetic vocabulary (everything on the device), a vocabulary with '##' word pieces through the
hybrid step (table path + host strings for the captions that contain a piece), and the same vocabulary with every
candidate through host strings (CONZIC_STRING_PATH=1, what the reference does each step).

    python tools/bench_vocab_paths.py [--sweeps 2] [--batch 64]
"""
import argparse
import json
import logging
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from conzic_b200 import gen_utils, runtime  # noqa: E402
from synthetic import synth  # noqa: E402
from conzic_b200.clip.clip import CLIP  # noqa: E402
from conzic_b200.models import BertMLM  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweeps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    B, n, K = a.batch, 10, 200
    log = logging.getLogger("paths")
    log.addHandler(logging.NullHandler())
    log.propagate = False
    bert_sd, clip_sd = synth.make_bert_state_dict(0), synth.make_clip_state_dict(0)
    pix = torch.stack([synth.make_pixel_values(i) for i in range(B)]).pin_memory()
    names = [f"img{i}.jpg" for i in range(B)]
    import tempfile
    hf_bert, hf_clip = synth.make_hf_tokenizers(tempfile.mkdtemp())
    for label, pieces, env in (("device (no pieces in the vocabulary)", False, {}),
                               ("hybrid (pieces: table path + host strings for flagged captions)", True, {}),
                               ("strings (pieces: every candidate through host strings)", True, {"CONZIC_STRING_PATH": "1"}),
                               ("hybrid, real transformers tokenizer classes (pieces, ~5 CLIP tokens per word)", "hf", {}),
                               ("strings, real transformers tokenizer classes", "hf", {"CONZIC_STRING_PATH": "1"})):
        os.environ.update(env)
        runtime.clear()
        bert = BertMLM(bert_sd)
        if pieces == "hf":
            ctok, btok = hf_clip, hf_bert
        else:
            ctok = synth.PieceCLIPTokenizer() if pieces else synth.SynthCLIPTokenizer()
            btok = synth.PieceBertTokenizer() if pieces else synth.SynthBertTokenizer()
        clip = CLIP(state_dict=clip_sd, tokenizer=ctok, processor=synth.SynthProcessor()).to("cuda:0")
        prompt = synth.hf_prompt() if pieces == "hf" else synth.SYNTH_PROMPT

        def call():
            return gen_utils.generate_caption(names, bert, clip, btok, pix, synth.make_token_mask("cuda"), log,
                                              prompt=prompt, batch_size=B, max_len=n, top_k=K,
                                              temperature=0.1, max_iter=a.sweeps, alpha=0.02, beta=2.0,
                                              generate_order="sequential")
        call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        texts, _ = call()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        merged = sum(1 for c in texts[-2] for w in c.split() if "p" in w[1:]) if pieces != "hf" else None
        print(json.dumps({"path": label, "batch": B, "sweeps": a.sweeps, "seconds": round(dt, 3),
                          "ms_per_gibbs_step": round(1e3 * dt / (a.sweeps * n), 2),
                          "merged_words_in_final_captions": merged}), flush=True)
        for k in env:
            os.environ.pop(k)
    runtime.clear()


if __name__ == "__main__":
    main()
