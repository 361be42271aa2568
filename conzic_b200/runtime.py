"""Binds the caller's (model, clip, tokenizer) objects -- duck-typed exactly as loosely as the reference
treats them -- to one libconzic context per weight pair, and caches it."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch

from . import synth, tokens
from .engine import Engine

_engines: Dict[Tuple[int, int, str], Engine] = {}
_last: Optional[Engine] = None
_tables: Dict[Tuple[int, int], tuple] = {}


def default_precision() -> str:
    """bf16 operands by default; CONZIC_PRECISION=bf16x3 selects the 3-pass split mode used for id parity."""
    return os.environ.get("CONZIC_PRECISION", "bf16")


def _dummy_bert_sd():
    H, V = 64, 8
    z = lambda *s: torch.zeros(*s)
    return {"bert.embeddings.word_embeddings.weight": z(V, H), "bert.embeddings.position_embeddings.weight": z(8, H),
            "bert.embeddings.token_type_embeddings.weight": z(2, H), "bert.embeddings.LayerNorm.weight": z(H),
            "bert.embeddings.LayerNorm.bias": z(H), "cls.predictions.bias": z(V),
            "cls.predictions.transform.dense.weight": z(H, H), "cls.predictions.transform.dense.bias": z(H),
            "cls.predictions.transform.LayerNorm.weight": z(H), "cls.predictions.transform.LayerNorm.bias": z(H),
            "bert.encoder.layer.0.intermediate.dense.weight": z(H, H)}


def engine_for(model, clip, tokenizer=None, precision: Optional[str] = None, device=None) -> Engine:
    """The engine holding `model`'s BERT weights and `clip`'s text tower.  Either may be None when only the
    other side is needed (the missing side is replaced by a 0-layer placeholder, or by the last engine that
    already holds the side that is present)."""
    global _last
    precision = precision or default_precision()
    if model is None or clip is None:
        for (mid, cid, pr), e in _engines.items():
            if pr == precision and ((model is None and cid == id(clip)) or (clip is None and mid == id(model))):
                return e
    key = (id(model), id(clip), precision)
    if key in _engines:
        return _engines[key]
    if clip is None:
        raise RuntimeError("no engine holds this model yet; call generate_caption or engine_for(model, clip) first")
    bert_sd = model.state_dict() if model is not None else _dummy_bert_sd()
    dev = device or getattr(clip, "device", None) or "cuda:0"
    if torch.device(dev).type != "cuda":
        dev = "cuda:0"
    kw = {}
    if tokenizer is not None:
        sp = [getattr(tokenizer, n, d) for n, d in (("pad_token_id", synth.PAD_ID), ("unk_token_id", synth.UNK_ID),
                                                    ("cls_token_id", synth.CLS_ID), ("sep_token_id", synth.SEP_ID),
                                                    ("mask_token_id", synth.MASK_ID))]
        kw["special_ids"] = [synth.SPECIAL_IDS[i] if s is None else int(s) for i, s in enumerate(sp)]
        kw["dot_id"] = int(tokenizer.vocab["."])
    eng = Engine(bert_sd, clip.state_dict(), device=dev, precision=precision, **kw)
    if tokenizer is not None and model is not None:
        bind_tokenizers(eng, tokenizer, clip.tokenizer)
    _engines[key] = eng
    _last = eng
    return eng


def bind_tokenizers(eng: Engine, bert_tokenizer, clip_tokenizer):
    """Builds (once per tokenizer pair) and uploads the BERT-id -> CLIP-id table."""
    key = (id(bert_tokenizer), id(clip_tokenizer))
    if key not in _tables:
        if type(bert_tokenizer) is synth.SynthBertTokenizer and type(clip_tokenizer) is synth.SynthCLIPTokenizer:
            off, tok = synth.build_bert2clip_table(clip_tokenizer.multi)
            needs_host = []
        else:
            off, tok, needs_host = tokens.build_bert2clip(bert_tokenizer, clip_tokenizer, eng.V,
                                                          [eng.cfg.pad_id, eng.cfg.unk_id, eng.cfg.cls_id,
                                                           eng.cfg.sep_id, eng.cfg.mask_id])
        _tables[key] = (off, tok, needs_host)
    off, tok, needs_host = _tables[key]
    eng.set_bert2clip(off, tok)
    eng.needs_host_ids = needs_host
    # host copies for the hybrid path of vocabularies with '##' pieces (tokens.plan_hybrid)
    eng.piece_mask_h = torch.zeros(eng.V, dtype=torch.bool)
    if needs_host:
        eng.piece_mask_h[torch.tensor(needs_host, dtype=torch.long)] = True
    eng.tok_len_h = (off[1:] - off[:-1])[: eng.V].to(torch.int32)


def any_engine() -> Engine:
    if _last is None:
        raise RuntimeError("no libconzic context exists yet")
    return _last


def clear():
    global _last
    for e in _engines.values():
        e.close()
    _engines.clear()
    _last = None
