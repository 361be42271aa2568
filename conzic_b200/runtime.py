"""Binds the caller's (model, clip, tokenizer) objects -- duck-typed exactly as loosely as the reference
treats them -- to one libconzic context per (weights, tokenizer, device, precision), and caches it."""
from __future__ import annotations

import os
import weakref
from typing import Dict, Optional, Tuple

import torch

from . import defaults, tokens
from .engine import Engine

_engines: Dict[tuple, Engine] = {}
_last: Optional[Engine] = None
_tables: Dict[Tuple[int, int], tuple] = {}


def default_precision() -> str:
    """"certified" unless CONZIC_PRECISION says otherwise (read when an engine is created):
    certified -- BERT / image tower in bf16x3, CLIP text tower in bf16 + exact re-score of every candidate the bf16
                 scores cannot rule out: the reference's token ids at close to bf16 speed;
    bf16x3    -- everything in the 3-pass split mode (fp32-grade; what `certified` is checked against);
    bf16      -- everything in bf16 (tolerance-only parity)."""
    return os.environ.get("CONZIC_PRECISION", "certified")


def resolve_device(dev) -> torch.device:
    """`clip.to("cuda")` in the reference gives an index-less device; the engine needs a concrete one."""
    d = torch.device(dev) if dev is not None else torch.device("cuda")
    if d.type != "cuda":
        d = torch.device("cuda")
    if d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def _dummy_bert_sd():
    H, V = 64, 8
    z = lambda *s: torch.zeros(*s)
    return {"bert.embeddings.word_embeddings.weight": z(V, H), "bert.embeddings.position_embeddings.weight": z(8, H),
            "bert.embeddings.token_type_embeddings.weight": z(2, H), "bert.embeddings.LayerNorm.weight": z(H),
            "bert.embeddings.LayerNorm.bias": z(H), "cls.predictions.bias": z(V),
            "cls.predictions.transform.dense.weight": z(H, H), "cls.predictions.transform.dense.bias": z(H),
            "cls.predictions.transform.LayerNorm.weight": z(H), "cls.predictions.transform.LayerNorm.bias": z(H),
            "bert.encoder.layer.0.intermediate.dense.weight": z(H, H)}


def _evict(key):
    global _last
    eng = _engines.pop(key, None)
    if eng is not None:
        if _last is eng:
            _last = None
        eng.close()


def engine_for(model, clip, tokenizer=None, precision: Optional[str] = None, device=None) -> Engine:
    """The engine holding `model`'s BERT weights and `clip`'s text tower.  Either may be None when only the
    other side is needed (the missing side is replaced by a 0-layer placeholder, or by the last engine that
    already holds the side that is present).  Engines are evicted when their model / clip object is collected."""
    global _last
    precision = precision or default_precision()
    dev = resolve_device(device or getattr(clip, "device", None))
    if model is None or clip is None:
        for (mid, cid, tid, pr, dv), e in _engines.items():
            if pr == precision and dv == str(dev) and ((model is None and cid == id(clip)) or
                                                       (clip is None and mid == id(model))):
                return e
    tid = id(tokenizer) if tokenizer is not None else 0
    key = (id(model), id(clip), tid, precision, str(dev))
    if key in _engines:
        return _engines[key]
    if clip is None:
        raise RuntimeError("no engine holds this model yet; call generate_caption or engine_for(model, clip) first")
    if tokenizer is None:  # any engine of these weights will do for tokenizer-free calls (CLIP-only methods)
        for k, e in _engines.items():
            if k[0] == id(model) and k[1] == id(clip) and k[3] == precision and k[4] == str(dev):
                return e
    bert_sd = model.state_dict() if model is not None else _dummy_bert_sd()
    kw = {}
    if tokenizer is not None:
        sp = [getattr(tokenizer, n, d) for n, d in (("pad_token_id", defaults.PAD_ID), ("unk_token_id", defaults.UNK_ID),
                                                    ("cls_token_id", defaults.CLS_ID), ("sep_token_id", defaults.SEP_ID),
                                                    ("mask_token_id", defaults.MASK_ID))]
        kw["special_ids"] = [defaults.SPECIAL_IDS[i] if s is None else int(s) for i, s in enumerate(sp)]
        kw["dot_id"] = int(tokenizer.vocab["."])
    eng = Engine(bert_sd, clip.state_dict(), device=dev, precision=precision, **kw)
    if tokenizer is not None and model is not None:
        bind_tokenizers(eng, tokenizer, clip.tokenizer)
    _engines[key] = eng
    _last = eng
    for obj in (model, clip, tokenizer):  # id() values are reused after collection: drop the engine with its owners
        if obj is not None:
            try:
                weakref.finalize(obj, _evict, key)
            except TypeError:
                pass
    return eng


def bind_tokenizers(eng: Engine, bert_tokenizer, clip_tokenizer):
    """Builds (once per tokenizer pair) and uploads the BERT-id -> CLIP-id table and, for the Hugging Face fast
    tokenizer pair, the device text pipeline's tables (vocabularies with '##' word pieces stay on the device)."""
    key = (id(bert_tokenizer), id(clip_tokenizer))
    special = [eng.cfg.pad_id, eng.cfg.unk_id, eng.cfg.cls_id, eng.cfg.sep_id, eng.cfg.mask_id]
    if key not in _tables:
        off, tok, pieces = tokens.build_bert2clip(bert_tokenizer, clip_tokenizer, eng.V, special)
        ok, why = tokens.text_vocab_supported(bert_tokenizer, clip_tokenizer)
        tv = tokens.build_text_vocab(bert_tokenizer, clip_tokenizer, eng.V, special, off, tok) if ok else None
        _tables[key] = (off, tok, pieces, tv, why)
        for obj in (bert_tokenizer, clip_tokenizer):  # id() values are reused after collection
            try:
                weakref.finalize(obj, _tables.pop, key, None)
            except TypeError:
                pass
    off, tok, pieces, tv, why = _tables[key]
    eng.set_bert2clip(off, tok)
    if tv is not None:
        eng.set_text_vocab(tv)
    # a vocabulary with '##' pieces that the device pipeline does not restate (duck-typed tokenizers) takes the
    # reference's string round trip every step
    eng.needs_strings = bool(pieces) and tv is None
    eng.text_vocab_note = why


def any_engine() -> Engine:
    if _last is None:
        raise RuntimeError("no libconzic context exists yet")
    return _last


def clear():
    global _last
    for e in list(_engines.values()):
        e.close()
    _engines.clear()
    _tables.clear()
    _last = None
