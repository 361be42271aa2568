#!/usr/bin/env python
"""Generate tests/golden/*.pt by running the UNMODIFIED reference in this container.

    python tools/make_golden.py            # needs /root/reference and transformers; CPU only

The reference modules (``gen_utils``, ``control_gen_utils``, ``clip.clip``, ``utils``) are imported
from /root/reference exactly as they are.  What is supplied around them:

* a 1-line ``colorlog`` shim, a stub ``sentiments_classifer`` module, and stub ``nltk.pos_tag`` /
  ``nltk.tokenize.word_tokenize`` callables under the unmodified ``POS_classifier``
  (NLTK / colorlog are not installed and there is no network);
* HF ``BertForMaskedLM(BertConfig())`` / ``CLIPModel(CLIPConfig())`` loaded with the synthetic
  state dicts from ``synthetic.synth`` (no pretrained weights exist offline);
* the synthetic string tokenizers of ``synthetic.synth``.

Recording is done by wrapping callables at run time (the model's forward, ``generate_caption_step``
and ``compute_image_text_similarity_via_raw_text``); no reference file is edited or copied.
Each fixture holds one record per Gibbs step plus the call's return value, and the crc32 of the
weights it was made with.  The GPU box never runs this script; it only reads the fixtures.
"""
import logging
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthetic import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
LOGIT_COLS = torch.arange(0, synth.BERT_VOCAB, 61)  # 501 strided vocabulary columns kept per row


def import_reference(sentiment_table):
    from oracle import ref_loader
    tok = synth.SynthBertTokenizer()
    return ref_loader.load(REF, sentiment_table, lambda w: tok.vocab[w], synth.synth_pos_tagger)


def build_models(bert_sd, clip_sd, CLIP, multi, pieces=False):
    from oracle import ref_loader
    ctok = synth.PieceCLIPTokenizer(multi) if pieces else synth.SynthCLIPTokenizer(multi)
    return ref_loader.build_models(bert_sd, clip_sd, CLIP, ctok, synth.SynthProcessor())


class Recorder:
    def __init__(self, bert, clip, mod, embed_stride=1):
        self.steps, self.cur, self.embed_stride = [], None, embed_stride
        self.bert, self.clip, self.mod = bert, clip, mod
        orig_fwd, orig_step = bert.forward, mod.generate_caption_step
        orig_sim = clip.compute_image_text_similarity_via_raw_text
        rec = self

        def fwd(inp, *a, **k):
            out = orig_fwd(inp, *a, **k)
            rec.fw = {"inp": inp.clone(), "logits": out.logits, "uses": 0}
            return out

        def step(out, gen_idx, mask, temperature=None, top_k=100):
            probs, ids = orig_step(out, gen_idx=gen_idx, mask=mask, temperature=temperature, top_k=top_k)
            # span order scores two positions from ONE forward: the second step's logits are stale w.r.t. the
            # token chosen at the first; such records are marked and skipped by the teacher-forced replays
            rec.cur = {"inp": rec.fw["inp"], "stale_logits": rec.fw["uses"] > 0}
            rec.fw["uses"] += 1
            row = rec.fw["logits"][:, gen_idx]
            cols = torch.cat([LOGIT_COLS.expand(row.shape[0], -1), ids], dim=1)
            rec.cur.update(pos=int(gen_idx), token_mask_dot=float(mask[0, synth.DOT_ID]),
                           logit_cols=cols.to(torch.int32), logit_vals=row.gather(1, cols).clone(),
                           logit_max=row.max(dim=1).values.clone(),
                           logit_lse=torch.logsumexp(row.double() / temperature, dim=1).float(),
                           probs=probs.clone(), idxs=ids.clone())
            return probs, ids

        def sim(image_embeds, text_list):
            ids = clip.tokenizer(text_list, padding=True, return_tensors="pt", max_length=77,
                                 truncation=True)["input_ids"]
            text_embeds = clip.compute_text_representation(text_list)
            score, ref = clip.compute_image_text_similarity_via_embeddings(image_embeds, text_embeds)
            keep = torch.arange(0, len(text_list), rec.embed_stride)
            rec.cur.update(clip_ids=ids.to(torch.int32), text_embeds=text_embeds[keep].clone(), embed_rows=keep,
                           clip_score=score.clone(), clip_ref=ref.clone(), n_texts=len(text_list),
                           image_embeds=image_embeds.clone())
            rec.steps.append(rec.cur)
            return score, ref

        bert.forward = fwd
        mod.generate_caption_step = step
        clip.compute_image_text_similarity_via_raw_text = sim
        self._restore = lambda: (setattr(mod, "generate_caption_step", orig_step))

    def close(self):
        self._restore()


CASES = [
    # name, kwargs
    dict(name="seq_b2_n4_k8", order="sequential", B=2, n=4, K=8, iters=2),
    dict(name="shuffle_b3_n5_k16_multi", order="shuffle", B=3, n=5, K=16, iters=2, multi=True),
    dict(name="random_b2_n3_k8", order="random", B=2, n=3, K=8, iters=2),
    dict(name="senti_seq_b2_n4_k8", order="sequential", B=2, n=4, K=8, iters=2, gamma=5.0, style="positive"),
    dict(name="senti_shuffle_neg_b2_n4_k8", order="shuffle", B=2, n=4, K=8, iters=2, gamma=5.0, style="negative"),
    dict(name="peaked_seq_b2_n4_k32", order="sequential", B=2, n=4, K=32, iters=2, peaked=True),
    dict(name="seq_b1_n10_k200", order="sequential", B=1, n=10, K=200, iters=1),
    dict(name="span_b2_n5_k8", order="span", B=2, n=5, K=8, iters=2),
    dict(name="pieces_seq_b2_n5_k16", order="sequential", B=2, n=5, K=16, iters=3, pieces=True),
    dict(name="pieces_shuffle_b3_n6_k16_multi", order="shuffle", B=3, n=6, K=16, iters=2, pieces=True, multi=True),
    dict(name="hf_shuffle_b3_n5_k24", order="shuffle", B=3, n=5, K=24, iters=2, hf=True),
    dict(name="pos_seq_b2_n5_k16", order="sequential", B=2, n=5, K=16, iters=2, gamma=5.0, ctl="pos"),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    table = synth.make_sentiment_table()
    utils, gen_utils, control_gen_utils, CLIP = import_reference(table)
    logger = logging.getLogger("golden")
    logger.addHandler(logging.NullHandler())
    clip_sd = synth.make_clip_state_dict(0)
    sds = {False: synth.make_bert_state_dict(0), True: synth.make_bert_state_dict(0, peaked=True)}
    only = set(sys.argv[1:])
    for case in CASES:
        if only and case["name"] not in only:
            continue
        peaked, multi = case.get("peaked", False), case.get("multi", False)
        bert_sd = sds[peaked]
        pieces = case.get("pieces", False)
        bert, clip = build_models(bert_sd, clip_sd, CLIP, multi, pieces)
        bert_tok = synth.PieceBertTokenizer if pieces else synth.SynthBertTokenizer
        prompt = synth.SYNTH_PROMPT
        if case.get("hf"):  # the real transformers tokenizer classes over generated vocabulary files
            import tempfile
            hf_bert, hf_clip = synth.make_hf_tokenizers(tempfile.mkdtemp())
            clip.tokenizer = hf_clip
            bert_tok = lambda: hf_bert
            prompt = synth.hf_prompt()
        B, n, K = case["B"], case["n"], case["K"]
        pix = torch.stack([synth.make_pixel_values(i) for i in range(B)])
        token_mask = synth.make_token_mask()
        gamma = case.get("gamma")
        mod = control_gen_utils if gamma is not None else gen_utils
        rec = Recorder(bert, clip, mod, embed_stride=13 if K >= 100 else 1)
        utils.set_seed(42)
        names = [f"img{i}.jpg" for i in range(B)]
        kw = dict(prompt=prompt, batch_size=B, max_len=n, top_k=K, temperature=0.1,
                  max_iter=case["iters"], alpha=0.02, beta=2.0, generate_order=case["order"])
        with torch.no_grad():
            if gamma is None:
                texts, scores = gen_utils.generate_caption(names, bert, clip, bert_tok(), pix,
                                                           token_mask, logger, **kw)
            else:
                texts, scores = control_gen_utils.control_generate_caption(
                    names, bert, clip, bert_tok(), pix, token_mask, logger, gamma=gamma,
                    ctl_type=case.get("ctl", "sentiment"), style_type=case.get("style", "positive"),
                    pos_type=synth.SYNTH_POS_TEMPLATE, **kw)
        rec.close()
        fixture = dict(case=case, steps=rec.steps, texts=texts, scores=scores,
                       bert_crc=synth.state_dict_checksum(bert_sd), clip_crc=synth.state_dict_checksum(clip_sd),
                       torch=torch.__version__)
        path = os.path.join(OUT, case["name"] + ".pt")
        torch.save(fixture, path)
        print(f"{case['name']}: {len(rec.steps)} steps, {os.path.getsize(path)/1024:.0f} KiB; final={texts[-2]}")


if __name__ == "__main__":
    main()
