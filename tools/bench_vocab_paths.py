#!/usr/bin/env python
"""Milliseconds per Gibbs step of generate_caption (BASELINE config 2 sizes, host pixels in, strings out) for the
candidate-text paths: the per-token table (vocabularies of whole words), the device text pipeline (real transformers
BertTokenizer / CLIPTokenizer classes with '##' pieces: WordPiece decode + CLIP BPE on the device) and the
reference's string round trip (CONZIC_STRING_PATH=1, or a duck-typed tokenizer pair with pieces).

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
    hf1_bert, hf1_clip = synth.make_hf_tokenizers(tempfile.mkdtemp(), single_token_words=True)
    S1 = {"CONZIC_STRING_PATH": "1"}
    for label, kind, env in (("device table path (synthetic tokenizers, no pieces, 1 CLIP token per word)", "synth", {}),
                             ("device text pipeline (real transformers classes, '##' pieces, 1 CLIP token per whole word)", "hf1", {}),
                             ("string round trip, same tokenizers", "hf1", S1),
                             ("device text pipeline (real transformers classes, '##' pieces, ~4 CLIP tokens per word)", "hf", {}),
                             ("string round trip, same tokenizers", "hf", S1),
                             ("string round trip (duck-typed tokenizers with pieces: what the engine falls back to)", "pieces", {})):
        os.environ.update(env)
        runtime.clear()
        bert = BertMLM(bert_sd)
        pieces = kind
        if kind == "hf":
            ctok, btok = hf_clip, hf_bert
        elif kind == "hf1":
            ctok, btok = hf1_clip, hf1_bert
        else:
            ctok = synth.PieceCLIPTokenizer() if kind == "pieces" else synth.SynthCLIPTokenizer()
            btok = synth.PieceBertTokenizer() if kind == "pieces" else synth.SynthBertTokenizer()
        clip = CLIP(state_dict=clip_sd, tokenizer=ctok, processor=synth.SynthProcessor()).to("cuda:0")
        prompt = synth.hf_prompt() if kind in ("hf", "hf1") else synth.SYNTH_PROMPT

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
        eng = runtime.any_engine()
        eng.profile(True)
        call()
        ph = {k: round(v[0] / (a.sweeps * n), 3) for k, v in eng.profile_read_phases().items() if v[1]}
        eng.profile(False)
        print(json.dumps({"path": label, "phases_ms_per_step": ph, "precision": eng.precision, "batch": B, "sweeps": a.sweeps, "seconds": round(dt, 3),
                          "ms_per_gibbs_step": round(1e3 * dt / (a.sweeps * n), 2),
                          "device_text_pipeline": bool(eng.has_text_vocab), "string_path": bool(env) or bool(eng.needs_strings),
                          "max_clip_tokens_per_word": eng.max_tok_per_word}), flush=True)
        for k in env:
            os.environ.pop(k)
    runtime.clear()


if __name__ == "__main__":
    main()
