"""Drop-in for the reference's `gen_utils` generation API (gen_utils.py:33-49, 51-146, 197-242, 289-333):
same function names, arguments, defaults, return structure, log lines and RNG consumption.  The per-position
work -- BERT row logits, masked top-k, candidate CLIP encoding, cosine/softmax, score fuse, argmax and the
write-back -- is ONE call into libconzic.so (`conzic_gibbs_step`); the host only walks the visiting order and,
once per sweep, reads ids and scores back to build the strings the reference returns and logs."""
from __future__ import annotations

import os
import random
import time

import numpy as np
import torch

from . import runtime
from .utils import get_init_text

__all__ = ["generate_caption_step", "sequential_generation", "shuffle_generation", "random_generation",
           "span_generation", "parallel_generation", "generate_caption"]


def generate_caption_step(out, gen_idx, mask, temperature=None, top_k=100):
    """softmax(out[:, gen_idx] / temperature) * mask -> top_k (probs, ids); gen_utils.py:33-49.
    Equal probabilities are ordered by ascending vocabulary id (torch leaves that order unspecified)."""
    eng = runtime.any_engine()
    row = out[:, gen_idx].to(eng.device, torch.float32).contiguous()
    m = mask.to(eng.device, torch.float32).contiguous()
    return eng.topk_mask(row, m, 1.0 if temperature is None else temperature, top_k)


class _Chain:
    """State of one generate call: ids on the device, which positions hold a word (sizes the CLIP tiles),
    and per-step result slots that are read back once per flush."""

    def __init__(self, model, clip, tokenizer, image_instance, token_mask, prompt, max_len, batch_size, ctl):
        self.eng = runtime.engine_for(model, clip, tokenizer)
        eng = self.eng
        self.clip = clip
        # Candidate captions become CLIP ids on the device: through the per-token table (vocabularies whose tokens are
        # whole words) or through the device text pipeline (Hugging Face BertTokenizer / CLIPTokenizer: '##' pieces,
        # punctuation clean-up, byte-level BPE).  The reference's string round trip remains for what neither covers --
        # duck-typed tokenizers whose vocabulary has '##' pieces -- and on request (CONZIC_STRING_PATH=1).
        self.string_path = os.environ.get("CONZIC_STRING_PATH") == "1" or bool(getattr(eng, "needs_strings", False))
        self.tokenizer, self.max_len, self.B = tokenizer, max_len, batch_size
        self.seed_len = len(prompt.split()) + 1
        batch = get_init_text(tokenizer, prompt, max_len, batch_size)
        self.image_embeds = clip.compute_image_representation_from_image_instance(image_instance)
        self.image_embeds = self.image_embeds.to(eng.device, torch.float32).contiguous()
        self.inp = torch.tensor(batch).to(eng.device)
        special = {eng.cfg.pad_id, eng.cfg.unk_id, eng.cfg.cls_id, eng.cfg.sep_id, eng.cfg.mask_id}
        self.holds_word = [t not in special for t in batch[0]]
        self.user_mask = token_mask
        if token_mask.device == eng.device and token_mask.dtype == torch.float32 and token_mask.is_contiguous():
            self.mask = token_mask
        else:
            self.mask = token_mask.to(eng.device, torch.float32).contiguous()
        n = max_len
        self.clip_slots = torch.zeros((n, batch_size), dtype=torch.float32, device=eng.device)
        self.senti_slots = torch.zeros((n, batch_size), dtype=torch.float32, device=eng.device) if ctl else None
        self.inp_slots = None

    def step_via_strings(self, slot, ii, top_k, temperature, alpha, beta, gamma=None, senti_table=None, logits_in=None,
                         pos_scorer=None):
        """The same step with the reference's string round trip (gen_utils.py:66-81): candidate ids -> host ->
        tokenizer.batch_decode -> CLIP tokenizer -> device.  Every arithmetic piece is still a libconzic kernel;
        only the text handling runs on the host.  Used by POS-template control (the tagger needs every string), by
        duck-typed tokenizers with '##' pieces, and when CONZIC_STRING_PATH=1 asks for it."""
        eng, tok = self.eng, self.tokenizer
        pos = self.seed_len + ii
        T = 1.0 if temperature is None else temperature
        self.mask.view(-1)[eng.cfg.dot_id] = 1.0 if ii == self.max_len - 1 else 0.0  # utils.py:53-59
        self.inp[:, pos] = eng.mask_id
        row = logits_in[:, : eng.V] if logits_in is not None else eng.bert_mlm_row(self.inp, pos)
        probs, idxs = eng.topk_mask(row, self.mask, T, top_k)
        # ids * token_mask[ids] and the candidate id tensor (gen_utils.py:71-74), on the host with the strings
        idxs_h = idxs.cpu()
        ids_masked_h = (idxs_h * self.mask.view(-1).cpu()[idxs_h]).long()
        cand = self.inp.cpu().unsqueeze(1).repeat(1, top_k, 1)
        cand[:, :, pos] = ids_masked_h
        flat = cand.view(-1, cand.shape[-1])
        texts = tok.batch_decode(flat, skip_special_tokens=True)
        clip_ids = self.clip.tokenize_texts(texts).to(eng.device, torch.int32)
        text_embeds = eng.clip_text_encode(clip_ids)
        senti_raw = repeats = best = None
        if pos_scorer is not None:
            # POS template (control_gen_utils.py:164-168): softmax_K(score / 0.1), no repeat penalty; the winner's
            # index comes back so its tag sequence and raw score can be reported (:171-178)
            tags, raw = pos_scorer(texts)
            raw = raw.to(torch.float32).view(self.B, top_k)
            senti_raw = (raw / 0.1).to(eng.device).contiguous()
            best = torch.empty((self.B,), dtype=torch.int64, device=eng.device)
        elif gamma is not None:
            special = torch.tensor(sorted({eng.cfg.pad_id, eng.cfg.unk_id, eng.cfg.cls_id, eng.cfg.sep_id, eng.cfg.mask_id}))
            table = senti_table.cpu()
            vis = ~torch.isin(flat, special)
            senti_raw = (table[flat] * vis).sum(1).view(self.B, top_k).to(eng.device).contiguous()
            repeats = ((ids_masked_h[:, :, None] == cand).float().sum(2) - 1).to(eng.device).contiguous()
        eng.score_select(text_embeds, self.image_embeds, probs, ids_masked_h.to(eng.device), self.inp, pos, alpha, beta,
                         gamma=gamma, senti_raw=senti_raw, repeats=repeats, out_clip_ref=self.clip_slots[slot],
                         out_senti=self.senti_slots[slot] if self.senti_slots is not None else None, out_best=best,
                         clip_ids=clip_ids)
        if best is not None:
            win = best.cpu()
            self.pos_scores = raw.gather(1, win.view(-1, 1)).squeeze(-1).numpy().tolist()
            self.pos_tags = [tags[int(win[i]) + i * top_k] for i in range(self.B)]
        self.holds_word[pos] = True

    def step(self, slot, ii, top_k, temperature, alpha, beta, gamma=None, senti_table=None, logits_in=None,
             pos_scorer=None):
        if self.string_path or pos_scorer is not None:
            return self.step_via_strings(slot, ii, top_k, temperature, alpha, beta, gamma, senti_table, logits_in,
                                         pos_scorer)
        pos = self.seed_len + ii
        before = sum(self.holds_word[:pos])
        after = sum(self.holds_word[pos + 1:])
        self.eng.gibbs_step(self.inp, self.mask, self.image_embeds, pos, ii == self.max_len - 1, top_k,
                            1.0 if temperature is None else temperature, alpha, beta, before, after, gamma=gamma,
                            senti_table=senti_table, out_clip_ref=self.clip_slots[slot],
                            out_senti=self.senti_slots[slot] if self.senti_slots is not None else None,
                            logits_in=logits_in)
        self.holds_word[pos] = True

    def finish(self):
        if self.mask is not self.user_mask:  # keep the reference's in-place side effect on the caller's mask
            self.user_mask.copy_(self.mask.to(self.user_mask.device, self.user_mask.dtype).view_as(self.user_mask))


def _sweeps(name, img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger, max_len, top_k,
            temperature, alpha, beta, max_iters, batch_size, verbose, order, gamma=None, senti_table=None,
            pos_scorer=None):
    """Sequential / shuffled sweeps (gen_utils.py:51-146, control_gen_utils.py:30-195).  `pos_scorer(texts) ->
    (tag sequences, f32[N] scores)` selects the POS-template formula, `senti_table` the sentiment one."""
    ch = _Chain(model, clip, tokenizer, image_instance, token_mask, prompt, max_len, batch_size, gamma is not None)
    best_score, best_caption = [0] * batch_size, ["None"] * batch_size
    texts, scores = [], []
    cur_text, cur_score = None, None
    for it in range(max_iters):
        for slot, ii in enumerate(order):
            ch.step(slot, ii, top_k, temperature, alpha, beta, gamma, senti_table, pos_scorer=pos_scorer)
        last = len(order) - 1
        ids = ch.inp.cpu()  # one device->host read per sweep
        cur_score = ch.clip_slots[last].cpu().numpy().tolist()
        cur_senti = ch.senti_slots[last].cpu().numpy().tolist() if gamma is not None else None
        if pos_scorer is not None:
            cur_senti = ch.pos_scores
        if verbose:
            shown = tokenizer.batch_decode(ids)
            cur_text = tokenizer.batch_decode(ids, skip_special_tokens=True)
            for jj in range(batch_size):
                if best_score[jj] < cur_score[jj]:
                    best_score[jj], best_caption[jj] = cur_score[jj], cur_text[jj]
                if gamma is None:
                    logger.info(f"iter {it + 1}, The {jj+1}-th image: {img_name[jj]},"
                                f"clip score {cur_score[jj]:.3f}: " + shown[jj])
                else:
                    logger.info(f"iter {it + 1}, The {jj+1}-th image: {img_name[jj]}, clip score {cur_score[jj]:.3f}"
                                f", ctl score {cur_senti[jj]:.3f}: " + shown[jj])
                    if pos_scorer is not None:
                        logger.info(ch.pos_tags[jj])
        texts.append(cur_text)
        scores.append(cur_score)
    texts.append(best_caption)
    scores.append(best_score)
    ch.finish()
    return texts, scores


def sequential_generation(img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger,
                          max_len=15, top_k=100, temperature=None, alpha=0.7, beta=1,
                          max_iters=20, batch_size=1, verbose=True):
    """One position at a time, left to right (gen_utils.py:51-96)."""
    return _sweeps("sequential", img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger, max_len,
                   top_k, temperature, alpha, beta, max_iters, batch_size, verbose, list(range(max_len)))


def shuffle_generation(img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger,
                       max_len=15, top_k=0, temperature=None, alpha=0.7, beta=1,
                       max_iters=20, batch_size=1, verbose=True):
    """One permutation per call from Python's global RNG, reused by every sweep (gen_utils.py:98-146)."""
    order = list(range(max_len))
    random.shuffle(order)
    logger.info(f"Order_list:{order}")
    return _sweeps("shuffle", img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger, max_len,
                   top_k, temperature, alpha, beta, max_iters, batch_size, verbose, order)


def random_generation(img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger,
                      max_len=15, top_k=0, temperature=None, alpha=0.7, beta=2, max_iters=300, print_every=10,
                      batch_size=1, verbose=True):
    """A uniformly random position per step from numpy's global RNG; best caption tracked after every step
    (gen_utils.py:197-242).  Steps are issued back to back; ids and scores of each step are kept on the
    device and read once per `print_every` steps."""
    ch = _Chain(model, clip, tokenizer, image_instance, token_mask, prompt, max_len, batch_size, False)
    eng = ch.eng
    group = max(1, int(print_every))
    ch.clip_slots = torch.zeros((group, batch_size), dtype=torch.float32, device=eng.device)
    inp_slots = torch.zeros((group,) + tuple(ch.inp.shape), dtype=torch.int64, device=eng.device)
    best_score, best_caption = [0] * batch_size, ["None"] * batch_size
    texts, scores = [], []
    done = 0
    while done < max_iters:
        n = min(group, max_iters - done)
        for slot in range(n):
            kk = np.random.randint(0, max_len)
            ch.step(slot, kk, top_k, temperature, alpha, beta)
            inp_slots[slot].copy_(ch.inp)
        ids_all = inp_slots[:n].cpu()
        sc_all = ch.clip_slots[:n].cpu().numpy().tolist()
        for slot in range(n):
            cur_text = tokenizer.batch_decode(ids_all[slot], skip_special_tokens=True)
            cur_score = sc_all[slot]
            for jj in range(batch_size):
                if best_score[jj] < cur_score[jj]:
                    best_score[jj], best_caption[jj] = cur_score[jj], cur_text[jj]
            step_no = done + slot + 1
            if verbose and step_no % print_every == 0:
                shown = tokenizer.batch_decode(ids_all[slot])
                for jj in range(batch_size):
                    logger.info(f"iter {step_no}, The {jj+1}-th image: {img_name[jj]},"
                                f"clip score {cur_score[jj]:.3f}: " + shown[jj])
                texts.append(cur_text)
                scores.append(cur_score)
        done += n
    texts.append(best_caption)
    scores.append(best_score)
    ch.finish()
    return texts, scores


def span_generation(img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger,
                    max_len=15, top_k=0, temperature=None, alpha=0.7, beta=1,
                    max_iters=20, batch_size=1, verbose=True):
    """Spans of two positions, left to right (gen_utils.py:148-195): both positions of a span are masked, BERT
    runs once on that state, and each position is then scored from those logits (the second one from logits
    that do not see the word just chosen for the first, exactly like the reference)."""
    ch = _Chain(model, clip, tokenizer, image_instance, token_mask, prompt, max_len, batch_size, False)
    eng = ch.eng
    span_len = 2
    best_score, best_caption = [0] * batch_size, ["None"] * batch_size
    texts, scores = [], []
    cur_text = None
    for it in range(max_iters):
        slot = 0
        for span_start in range(0, max_len, span_len):
            span_end = min(span_start + span_len, max_len)
            lo, hi = ch.seed_len + span_start, ch.seed_len + span_end
            ch.inp[:, lo:hi] = eng.mask_id
            for p in range(lo, hi):
                ch.holds_word[p] = False
            rows = [eng.bert_mlm_row_padded(ch.inp, p) for p in range(lo, hi)]  # same ids -> same forward
            for ii in range(span_start, span_end):
                ch.step(slot, ii, top_k, temperature, alpha, beta, logits_in=rows[ii - span_start])
                slot += 1
        ids = ch.inp.cpu()
        cur_score = ch.clip_slots[slot - 1].cpu().numpy().tolist()
        if verbose:
            shown = tokenizer.batch_decode(ids)
            cur_text = tokenizer.batch_decode(ids, skip_special_tokens=True)
            for jj in range(batch_size):
                if best_score[jj] < cur_score[jj]:
                    best_score[jj], best_caption[jj] = cur_score[jj], cur_text[jj]
                logger.info(f"iter {it + 1}, The {jj+1}-th image: {img_name[jj]},"
                            f"clip score {cur_score[jj]:.3f}: " + shown[jj])
        texts.append(cur_text)
        scores.append(cur_score)
    texts.append(best_caption)
    scores.append(best_score)
    ch.finish()
    return texts, scores


def parallel_generation(*args, **kwargs):
    raise NotImplementedError("'parallel' is unreachable from the reference CLI and is not provided")


def generate_caption(img_name, model, clip, tokenizer, image_instance, token_mask, logger,
                     prompt="", batch_size=1, max_len=15,
                     top_k=100, temperature=1.0, max_iter=500, alpha=0.7, beta=1,
                     generate_order="sequential"):
    """Entry point used by demo.py / run.py (gen_utils.py:289-333): returns (generate_texts, clip_scores), one
    list per sweep plus the best-by-CLIP-score list last; [-2] is the final sweep."""
    start_time = time.time()
    common = dict(batch_size=batch_size, max_len=max_len, top_k=top_k, alpha=alpha, beta=beta,
                  temperature=temperature)
    if generate_order == "sequential":
        generate_texts, clip_scores = sequential_generation(img_name, model, clip, tokenizer, image_instance,
                                                            token_mask, prompt, logger, max_iters=max_iter, **common)
    elif generate_order == "shuffle":
        generate_texts, clip_scores = shuffle_generation(img_name, model, clip, tokenizer, image_instance, token_mask,
                                                         prompt, logger, max_iters=max_iter, **common)
    elif generate_order == "random":
        generate_texts, clip_scores = random_generation(img_name, model, clip, tokenizer, image_instance, token_mask,
                                                        prompt, logger, max_iters=max_iter * max_len,
                                                        print_every=max_len, verbose=True, **common)
    elif generate_order == "span":
        generate_texts, clip_scores = span_generation(img_name, model, clip, tokenizer, image_instance, token_mask,
                                                      prompt, logger, max_iters=max_iter, **common)
    elif generate_order == "parallel":
        return parallel_generation()
    else:
        raise ValueError(f"unknown generate_order {generate_order!r}")
    logger.info("Finished in %.3fs" % (time.time() - start_time))
    final_caption, best_caption = generate_texts[-2], generate_texts[-1]
    for i in range(batch_size):
        logger.info(f"The {i+1}-th image: {img_name[i]}")
        logger.info(f"final caption: {final_caption[i]}")
        logger.info(f"best caption: {best_caption[i]}")
    return generate_texts, clip_scores
