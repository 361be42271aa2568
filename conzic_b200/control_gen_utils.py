"""Drop-in for the sentiment-controlled half of the reference's `control_gen_utils`
(control_gen_utils.py:30-134, 197-232): score = alpha*p_LM + beta*p_CLIP + gamma*softmax_K(control) +
0.1*(1-exp(repeats)).

The reference scores every candidate caption with NLTK/SentiWordNet on the CPU (sentiments_classifer.py).
Here the control signal is a per-vocabulary-id table f32[V] evaluated on the device (candidates of one row
differ in one word and the softmax over K is shift invariant); building that table from SentiWordNet is host
tooling outside this path.  Register one with `set_sentiment_table` or pass
`sentiment_table=`.

POS-template control (control_gen_utils.py:136-195, POS_classifier.py:6-31) scores whole candidate captions with a
part-of-speech tagger, which is host string code: `POS_sequential_generation` therefore runs the step through the
string path (BERT, top-k, CLIP encode and the fused score / argmax stay libconzic kernels) and takes the tagger as
a plug-in -- `set_pos_tagger(fn)` with `fn(text) -> list of universal tags`; the default loads NLTK lazily."""
from __future__ import annotations

import random
import time

import torch

from . import runtime
from .gen_utils import _sweeps, generate_caption_step  # noqa: F401  (re-exported like the reference's duplicate)

_default_table = None


def set_sentiment_table(table):
    """f32[V]: control score of each BERT vocabulary id for the "positive" style ("negative" negates it)."""
    global _default_table
    _default_table = table


def _table_for(model, clip, tokenizer, ctl_signal, sentiment_table):
    t = sentiment_table if sentiment_table is not None else _default_table
    if t is None:
        raise RuntimeError("sentiment control needs a per-vocabulary score table: call "
                           "conzic_b200.control_gen_utils.set_sentiment_table(table) or pass sentiment_table=")
    eng = runtime.engine_for(model, clip, tokenizer)
    t = torch.as_tensor(t, dtype=torch.float32)
    if ctl_signal == "negative":  # sentiments_classifer.py:31-32
        t = -t
    return t.to(eng.device).contiguous()


def sentiment_sequential_generation(img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger,
                                    max_len=15, top_k=0, temperature=None, alpha=0.7, beta=1,
                                    max_iters=20, batch_size=1,
                                    verbose=True, gamma=5, ctl_signal="positive", sentiment_table=None):
    table = _table_for(model, clip, tokenizer, ctl_signal, sentiment_table)
    return _sweeps("sentiment_sequential", img_name, model, clip, tokenizer, image_instance, token_mask, prompt,
                   logger, max_len, top_k, temperature, alpha, beta, max_iters, batch_size, verbose,
                   list(range(max_len)), gamma=gamma, senti_table=table)


def sentiment_shuffle_generation(img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger,
                                 max_len=15, top_k=0, temperature=None, alpha=0.7, beta=1,
                                 max_iters=20, batch_size=1,
                                 verbose=True, gamma=5, ctl_signal="positive", sentiment_table=None):
    table = _table_for(model, clip, tokenizer, ctl_signal, sentiment_table)
    order = list(range(max_len))
    random.shuffle(order)
    logger.info(f"Order_list:{order}")
    return _sweeps("sentiment_shuffle", img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger,
                   max_len, top_k, temperature, alpha, beta, max_iters, batch_size, verbose, order, gamma=gamma,
                   senti_table=table)


_pos_tagger = None


def set_pos_tagger(fn):
    """fn(text) -> list[str]: universal part-of-speech tags of the word tokens of `text`
    (what `nltk.pos_tag(word_tokenize(text), tagset="universal")` yields at POS_classifier.py:12-14)."""
    global _pos_tagger
    _pos_tagger = fn


def _nltk_tagger(text):
    try:
        from nltk import pos_tag
        from nltk.tokenize import word_tokenize
    except ImportError as e:  # not installable offline; the caller supplies a tagger instead
        raise RuntimeError("POS control needs a tagger: install nltk (+ punkt, averaged_perceptron_tagger, "
                           "universal_tagset) or call conzic_b200.control_gen_utils.set_pos_tagger(fn)") from e
    return [t for _, t in pos_tag(word_tokenize(text), tagset="universal")]


def batch_texts_POS_analysis(batch_texts, pos_templete, device="cuda", tagger=None):
    """Share of template slots a caption's tag sequence satisfies (POS_classifier.py:6-31).  The tag sequence
    is cut or padded with "" to the template length; an empty template slot accepts anything; otherwise the
    tag must be `in` the slot (membership for a list slot, substring for a string slot, as in the reference).
    Returns (tag sequences, f32[N] scores on the host)."""
    tagger = tagger or _pos_tagger or _nltk_tagger
    n_slots = len(pos_templete)
    all_tags, scores = [], torch.zeros(len(batch_texts))
    for i, text in enumerate(batch_texts):
        tags = list(tagger(text))
        fitted = (tags + [""] * n_slots)[:n_slots]
        hits = sum(1 for slot, tag in zip(pos_templete, fitted) if slot == "" or tag in slot)
        all_tags.append(tags)
        scores[i] = hits / n_slots
    return all_tags, scores


def POS_sequential_generation(img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger,
                              max_len=15, top_k=0, temperature=None, alpha=0.7, beta=1, gamma=0.1,
                              max_iters=20, batch_size=1, ctl_signal=["DET"], verbose=True):
    """Left-to-right sweeps with alpha*p_LM + beta*p_CLIP + gamma*softmax_K(template score / 0.1)
    (control_gen_utils.py:136-195)."""
    logger.info(ctl_signal)
    return _sweeps("pos_sequential", img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger,
                   max_len, top_k, temperature, alpha, beta, max_iters, batch_size, verbose, list(range(max_len)),
                   gamma=gamma, pos_scorer=lambda texts: batch_texts_POS_analysis(texts, ctl_signal))


def control_generate_caption(img_name, model, clip, tokenizer, image_instance, token_mask, logger,
                             prompt="", batch_size=10, max_len=25,
                             top_k=100, temperature=1.0, max_iter=500, alpha=0.7, beta=1, gamma=5,
                             ctl_type="sentiment", style_type="positive", pos_type=None, generate_order="sequential",
                             sentiment_table=None):
    """control_gen_utils.py:197-232: sequential order -> sequential sweeps, any other order -> shuffled sweeps."""
    start_time = time.time()
    if ctl_type == "sentiment":
        fn = sentiment_sequential_generation if generate_order == "sequential" else sentiment_shuffle_generation
        generate_texts, clip_scores = fn(img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger,
                                         batch_size=batch_size, max_len=max_len, top_k=top_k, alpha=alpha, beta=beta,
                                         gamma=gamma, temperature=temperature, max_iters=max_iter,
                                         ctl_signal=style_type, sentiment_table=sentiment_table)
    else:  # POS control, always left to right (control_gen_utils.py:220-224)
        generate_texts, clip_scores = POS_sequential_generation(
            img_name, model, clip, tokenizer, image_instance, token_mask, prompt, logger, batch_size=batch_size,
            max_len=max_len, top_k=top_k, alpha=alpha, beta=beta, gamma=gamma, temperature=temperature,
            ctl_signal=pos_type, max_iters=max_iter)
    logger.info("Finished in %.3fs" % (time.time() - start_time))
    final_caption, best_caption = generate_texts[-2], generate_texts[-1]
    for i in range(batch_size):
        logger.info(f"The {i+1}-th image: {img_name[i]}")
        logger.info(f"final caption: {final_caption[i]}")
        logger.info(f"best caption: {best_caption[i]}")
    return generate_texts, clip_scores
