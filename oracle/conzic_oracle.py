"""CPU oracle for the ConZIC Gibbs-BERT caption-polishing path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import
this module; the product package ``conzic_b200`` never does (it fails loudly without its CUDA
library instead of falling back to anything here).

What this restates, in plain fp32 torch on the CPU (the reference's own arithmetic is fp32
eager torch, so torch CPU ops are the faithful medium):

* the reference's own loop code: ``gen_utils.py:33-49`` (``generate_caption_step``),
  ``gen_utils.py:51-96 / 98-146 / 197-242`` (sequential / shuffle / random order),
  ``control_gen_utils.py:30-134`` (sentiment variants), ``control_gen_utils.py:136-195`` +
  ``POS_classifier.py:6-31`` (POS-template variant; the NLTK tagger itself is a plug-in), ``utils.py:46-59`` (init text, '.' mask
  rule), ``clip/clip.py:48-102`` (image / text representation, similarity);
* the third-party arithmetic those call, which is NOT under ``/root/reference``:
  ``transformers`` (unpinned in ``requirements.txt:3``; 5.5.0 is what is installed here):
  ``BertForMaskedLM.forward`` (``models/bert/modeling_bert.py:72-112, 129-137, 179-181, 294-298,
  339-356, 481-501``) and ``CLIPModel.text_model / vision_model`` (``models/clip/modeling_clip.py:
  253-256, 271-298, 347-384, 531-589, 676-686``), with weights passed as HF state dicts.

Pinning: the reference repo has no tests or golden vectors for this path.
The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build
container: ``tools/make_golden.py`` imports the unmodified ``gen_utils`` / ``control_gen_utils``
/ ``clip.clip`` from ``/root/reference``, drives them with HF ``BertForMaskedLM`` / ``CLIPModel``
loaded with the synthetic state dicts, records every step, and stores the records under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through this file.
"""
from __future__ import annotations

import math
import random
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------
# BERT-base masked LM  (HF modeling_bert.py)
# ----------------------------------------------------------------------------------------
def bert_encoder(sd: SD, inp: torch.Tensor, heads: int = 12, eps: float = 1e-12) -> torch.Tensor:
    """ids i64[B,L] -> hidden f32[B,L,768].  No attention mask, token_type 0 (gen_utils.py:69)."""
    B, L = inp.shape
    e = "bert.embeddings."
    x = sd[e + "word_embeddings.weight"][inp]
    x = x + sd[e + "token_type_embeddings.weight"][0]
    x = x + sd[e + "position_embeddings.weight"][:L]
    H = x.shape[-1]
    x = F.layer_norm(x, (H,), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], eps)
    i = 0
    while f"bert.encoder.layer.{i}.attention.self.query.weight" in sd:
        p = f"bert.encoder.layer.{i}."
        q = F.linear(x, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"])
        k = F.linear(x, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"])
        v = F.linear(x, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"])
        sh = lambda t: t.view(B, L, heads, H // heads).transpose(1, 2)
        a = F.scaled_dot_product_attention(sh(q), sh(k), sh(v))
        a = a.transpose(1, 2).reshape(B, L, H)
        a = F.linear(a, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
        x = F.layer_norm(a + x, (H,), sd[p + "attention.output.LayerNorm.weight"],
                         sd[p + "attention.output.LayerNorm.bias"], eps)
        h = F.gelu(F.linear(x, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
        h = F.linear(h, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
        x = F.layer_norm(h + x, (H,), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], eps)
        i += 1
    return x


def bert_mlm_head(sd: SD, x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """hidden f32[...,768] -> logits f32[...,V]  (modeling_bert.py:481-501, decoder tied)."""
    p = "cls.predictions."
    H = x.shape[-1]
    h = F.gelu(F.linear(x, sd[p + "transform.dense.weight"], sd[p + "transform.dense.bias"]))
    h = F.layer_norm(h, (H,), sd[p + "transform.LayerNorm.weight"], sd[p + "transform.LayerNorm.bias"], eps)
    return F.linear(h, sd["bert.embeddings.word_embeddings.weight"], sd[p + "bias"])


def bert_mlm_logits(sd: SD, inp: torch.Tensor) -> torch.Tensor:
    """What ``model(inp).logits`` returns at gen_utils.py:69: logits for EVERY position."""
    return bert_mlm_head(sd, bert_encoder(sd, inp))


# ----------------------------------------------------------------------------------------
# CLIP ViT-B/32  (HF modeling_clip.py)
# ----------------------------------------------------------------------------------------
def _clip_layers(sd: SD, prefix: str, x: torch.Tensor, heads: int, causal: bool, eps: float = 1e-5):
    N, T, H = x.shape
    i = 0
    while f"{prefix}.encoder.layers.{i}.layer_norm1.weight" in sd:
        p = f"{prefix}.encoder.layers.{i}."
        h = F.layer_norm(x, (H,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"])
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"])
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
        sh = lambda t: t.view(N, T, heads, H // heads).transpose(1, 2)
        a = F.scaled_dot_product_attention(sh(q), sh(k), sh(v), is_causal=causal)
        a = a.transpose(1, 2).reshape(N, T, H)
        x = x + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (H,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps)
        h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
        h = h * torch.sigmoid(1.702 * h)  # quick_gelu, HF activations.py:122-123
        x = x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        i += 1
    return x


def clip_text_embeds(sd: SD, ids: torch.Tensor, eos_id: int = 49407) -> torch.Tensor:
    """CLIP ids i64[N,T] (right-padded with EOS) -> projected text embeds f32[N,512]
    (clip/clip.py:78-83).  Causal tower pooled at the first EOS, so the pad mask is inert."""
    N, T = ids.shape
    e = "text_model.embeddings."
    x = sd[e + "token_embedding.weight"][ids] + sd[e + "position_embedding.weight"][:T]
    x = _clip_layers(sd, "text_model", x, heads=8, causal=True)
    H = x.shape[-1]
    x = F.layer_norm(x, (H,), sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"], 1e-5)
    first_eos = (ids == eos_id).int().argmax(dim=-1)
    pooled = x[torch.arange(N), first_eos]
    return F.linear(pooled, sd["text_projection.weight"])


def clip_image_embeds(sd: SD, pixel_values: torch.Tensor) -> torch.Tensor:
    """pixels f32[B,3,224,224] -> projected image embeds f32[B,512]  (clip/clip.py:59-61)."""
    e = "vision_model.embeddings."
    B = pixel_values.shape[0]
    pe = F.conv2d(pixel_values, sd[e + "patch_embedding.weight"], stride=32).flatten(2).transpose(1, 2)
    x = torch.cat([sd[e + "class_embedding"].expand(B, 1, -1), pe], dim=1) + sd[e + "position_embedding.weight"]
    H = x.shape[-1]
    x = F.layer_norm(x, (H,), sd["vision_model.pre_layrnorm.weight"], sd["vision_model.pre_layrnorm.bias"], 1e-5)
    x = _clip_layers(sd, "vision_model", x, heads=12, causal=False)
    pooled = F.layer_norm(x[:, 0], (H,), sd["vision_model.post_layernorm.weight"],
                          sd["vision_model.post_layernorm.bias"], 1e-5)
    return F.linear(pooled, sd["visual_projection.weight"])


def image_text_similarity(image_embeds: torch.Tensor, text_embeds: torch.Tensor, logit_scale: torch.Tensor):
    """clip/clip.py:86-98: returns (softmax over K, cosine), both f32[B,K]."""
    text_embeds = text_embeds.view(image_embeds.shape[0], -1, text_embeds.shape[-1])
    image_embeds = image_embeds / image_embeds.norm(dim=-1, keepdim=True)
    text_embeds = text_embeds / text_embeds.norm(dim=-1, keepdim=True)
    scale = logit_scale.exp()
    logits = torch.matmul(text_embeds, image_embeds.unsqueeze(-1)).squeeze(-1) * scale
    return logits.softmax(dim=1), logits / scale


# ----------------------------------------------------------------------------------------
# reference loop pieces
# ----------------------------------------------------------------------------------------
def generate_caption_step(out, gen_idx, mask, temperature=None, top_k=100):
    """gen_utils.py:33-49 (dup control_gen_utils.py:12-28)."""
    logits = out[:, gen_idx]
    if temperature is not None:
        logits = logits / temperature
    probs = F.softmax(logits, dim=-1)
    probs = probs * mask
    return probs.topk(top_k, dim=-1)


def get_init_text(tokenizer, seed_text, max_len, batch_size=1):
    """utils.py:46-51."""
    ids = tokenizer.encode(seed_text + tokenizer.mask_token * max_len)
    return [ids] * batch_size


def update_token_mask(tokenizer, token_mask, max_len, index):
    """utils.py:53-59: '.' allowed only at the last position; mutates in place."""
    token_mask[:, tokenizer.vocab["."]] = 1 if index == max_len - 1 else 0
    return token_mask


def table_sentiment(batch_texts: List[str], table: torch.Tensor, tokenizer, temperature, batch_size_image):
    """Stand-in for sentiments_classifer.py:35-48 with a per-word table instead of SentiWordNet
    (NLTK is not installable offline): score(text) = sum of table[word], softmax over K."""
    scores = torch.zeros(len(batch_texts))
    for i, t in enumerate(batch_texts):
        s = 0.0
        for w in t.split():
            s += float(table[tokenizer.vocab[w]])
        scores[i] = s
    scores = scores.view(batch_size_image, -1)
    return F.softmax(scores / temperature, dim=1), scores


def pos_template_scores(batch_texts: List[str], template, tagger):
    """POS_classifier.py:6-31: tag every caption, fit the tag list to the template length (cut, or pad with ""),
    count slots that are empty or contain the tag (Python ``in``), score = count / slots.  ``tagger(text)`` stands
    in for ``nltk.pos_tag(word_tokenize(text), tagset="universal")``."""
    scores = torch.zeros(len(batch_texts))
    tags_out = []
    for b, text in enumerate(batch_texts):
        res = list(tagger(text))
        total = len(template)
        cur = res + [""] * (total - len(res)) if len(res) <= total else res[:total]
        correct = 0
        for w in range(len(cur)):
            if template[w] == "" or cur[w] in template[w]:
                correct += 1
        tags_out.append(res)
        scores[b] = correct / total
    return tags_out, scores


class Oracle:
    """Holds the two state dicts and the tokenizers; methods mirror the reference's callables."""

    def __init__(self, bert_sd: SD, clip_sd: SD, tokenizer, clip_tokenizer, sentiment_table=None,
                 full_logits: bool = True):
        self.bert_sd, self.clip_sd = bert_sd, clip_sd
        self.tokenizer, self.clip_tokenizer = tokenizer, clip_tokenizer
        self.sentiment_table = sentiment_table
        # full_logits=True does what the reference does (MLM head on all L rows, gen_utils.py:69);
        # False evaluates the head on the one row that is used -- same numbers, less CPU time.
        self.full_logits = full_logits
        self.trace: Optional[list] = None
        self.timers = {"bert": 0.0, "clip_text": 0.0, "strings": 0.0}

    # -- clip/clip.py -------------------------------------------------------------------
    def compute_image_representation(self, pixel_values):
        return clip_image_embeds(self.clip_sd, pixel_values)

    def compute_text_representation(self, text_list):
        t0 = time.perf_counter()
        ids = self.clip_tokenizer(text_list, padding=True, return_tensors="pt",
                                  max_length=self.clip_tokenizer.max_len_single_sentence + 2,
                                  truncation=True)["input_ids"]
        t1 = time.perf_counter()
        out = clip_text_embeds(self.clip_sd, ids)
        self.timers["strings"] += t1 - t0
        self.timers["clip_text"] += time.perf_counter() - t1
        return out, ids

    # -- one Gibbs step: gen_utils.py:66-81 / control_gen_utils.py:45-67 --------------------
    def step(self, inp, image_embeds, token_mask, pos, ii, max_len, top_k, temperature, alpha, beta,
             gamma=None, ctl_signal="positive", logits_row=None, pos_template=None, pos_tagger=None):
        """`logits_row` (f32[B,V]): logits of row `pos` from an earlier forward (span order, gen_utils.py:162-165);
        None = run BERT on the current ids."""
        tok = self.tokenizer
        token_mask = update_token_mask(tok, token_mask, max_len, ii)
        inp[:, pos] = tok.mask_token_id
        inp_ = inp.clone()
        t0 = time.perf_counter()
        if logits_row is not None:
            probs, idxs = generate_caption_step(logits_row[:, None], 0, token_mask, temperature, top_k)
        elif self.full_logits:
            out = bert_mlm_logits(self.bert_sd, inp)
            probs, idxs = generate_caption_step(out, pos, token_mask, temperature, top_k)
            logits_row = out[:, pos]
        else:
            logits_row = bert_mlm_head(self.bert_sd, bert_encoder(self.bert_sd, inp)[:, pos])
            probs, idxs = generate_caption_step(logits_row[:, None], 0, token_mask, temperature, top_k)
        self.timers["bert"] += time.perf_counter() - t0
        topk_inp = inp_.unsqueeze(1).repeat(1, top_k, 1)
        idxs_ = (idxs * token_mask[0][idxs]).long()
        topk_inp[:, :, pos] = idxs_
        t0 = time.perf_counter()
        texts = tok.batch_decode(topk_inp.view(-1, topk_inp.shape[-1]), skip_special_tokens=True)
        self.timers["strings"] += time.perf_counter() - t0
        text_embeds, clip_ids = self.compute_text_representation(texts)
        clip_score, clip_ref = image_text_similarity(image_embeds, text_embeds, self.clip_sd["logit_scale"])
        final = alpha * probs + beta * clip_score
        senti_scores = None
        if pos_template is not None:  # control_gen_utils.py:164-168
            _, pos_scores = pos_template_scores(texts, pos_template, pos_tagger)
            senti_scores = pos_scores.view(inp.shape[0], -1)
            final = final + gamma * torch.softmax(senti_scores / 0.1, dim=-1)
        elif gamma is not None:
            repeats = (idxs_[:, :, None] == topk_inp).float().sum(2) - 1
            table = -self.sentiment_table if ctl_signal == "negative" else self.sentiment_table
            senti_probs, senti_scores = table_sentiment(texts, table, tok, 1, inp.shape[0])
            final = final + gamma * senti_probs + 0.1 * (1 - torch.exp(repeats))
        best = final.argmax(dim=1).view(-1, 1)
        inp[:, pos] = idxs_.gather(1, best).squeeze(-1)
        cur_clip = clip_ref.gather(1, best).squeeze(-1)
        if self.trace is not None:
            self.trace.append(dict(pos=pos, ii=ii, inp_before=inp_.clone(), logits_row=logits_row.clone(),
                                   probs=probs.clone(), idxs=idxs.clone(), idxs_masked=idxs_.clone(),
                                   clip_ids=clip_ids.clone(), text_embeds=text_embeds.clone(),
                                   clip_score=clip_score.clone(), clip_ref=clip_ref.clone(),
                                   final=final.clone(), best=best.view(-1).clone(), inp_after=inp.clone()))
        senti = senti_scores.gather(1, best).squeeze(-1).tolist() if senti_scores is not None else None
        return cur_clip.tolist(), senti

    # -- the loops: gen_utils.py:51-146,197-242; control_gen_utils.py:30-134 ---------------------
    def generate(self, pixel_values, token_mask, prompt, order="sequential", max_len=10, top_k=200,
                 temperature=0.1, alpha=0.02, beta=2.0, max_iters=5, gamma=None, ctl_signal="positive",
                 image_embeds=None, pos_template=None, pos_tagger=None):
        """Returns (gen_texts_list, clip_score_sequence) with the reference's structure:
        one list per sweep plus the best-by-CLIP-score list last (gen_utils.py:93-96)."""
        tok = self.tokenizer
        B = pixel_values.shape[0] if image_embeds is None else image_embeds.shape[0]
        seed_len = len(prompt.split()) + 1
        batch = get_init_text(tok, prompt, max_len, B)
        if image_embeds is None:
            image_embeds = self.compute_image_representation(pixel_values)
        inp = torch.tensor(batch)
        best_score, best_cap = [0] * B, ["None"] * B
        texts_list, scores_list = [], []
        if order == "sequential":
            positions = list(range(max_len))
        elif order == "shuffle" or (gamma is not None and order != "sequential"):
            positions = list(range(max_len))
            random.shuffle(positions)  # gen_utils.py:110-111, one permutation per call
        if order == "span" and gamma is None:
            # gen_utils.py:148-195: spans of 2 positions are masked together and scored from ONE forward
            for _ in range(max_iters):
                for span_start in range(0, max_len, 2):
                    span_end = min(span_start + 2, max_len)
                    inp[:, seed_len + span_start: seed_len + span_end] = tok.mask_token_id
                    rows = bert_mlm_head(self.bert_sd, bert_encoder(self.bert_sd, inp)[:, seed_len + span_start: seed_len + span_end])
                    for ii in range(span_start, span_end):
                        cur, _ = self.step(inp, image_embeds, token_mask, seed_len + ii, ii, max_len, top_k, temperature,
                                           alpha, beta, logits_row=rows[:, ii - span_start])
                cur_text = tok.batch_decode(inp, skip_special_tokens=True)
                for jj in range(B):
                    if best_score[jj] < cur[jj]:
                        best_score[jj], best_cap[jj] = cur[jj], cur_text[jj]
                texts_list.append(cur_text)
                scores_list.append(cur)
        elif order == "random" and gamma is None:
            # gen_utils.py:197-242 with generate_caption's max_iter*=max_len, print_every=max_len
            for step in range(max_iters * max_len):
                kk = np.random.randint(0, max_len)
                cur, _ = self.step(inp, image_embeds, token_mask, seed_len + kk, kk, max_len, top_k,
                                   temperature, alpha, beta)
                cur_text = tok.batch_decode(inp, skip_special_tokens=True)
                for jj in range(B):
                    if best_score[jj] < cur[jj]:
                        best_score[jj], best_cap[jj] = cur[jj], cur_text[jj]
                if (step + 1) % max_len == 0:
                    texts_list.append(cur_text)
                    scores_list.append(cur)
        else:
            for _ in range(max_iters):
                for ii in positions:
                    cur, _ = self.step(inp, image_embeds, token_mask, seed_len + ii, ii, max_len, top_k,
                                       temperature, alpha, beta, gamma, ctl_signal, pos_template=pos_template,
                                       pos_tagger=pos_tagger)
                cur_text = tok.batch_decode(inp, skip_special_tokens=True)
                for jj in range(B):
                    if best_score[jj] < cur[jj]:
                        best_score[jj], best_cap[jj] = cur[jj], cur_text[jj]
                texts_list.append(cur_text)
                scores_list.append(cur)
        texts_list.append(best_cap)
        scores_list.append(best_score)
        self.final_ids = inp
        return texts_list, scores_list
