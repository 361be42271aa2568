"""BERT-id -> CLIP-BPE-id table: the device-side replacement of the reference's per-step string round
trip `tokenizer.batch_decode(...)` -> `CLIPTokenizer(...)` (gen_utils.py:75, clip/clip.py:71-72).

For a word-level BERT token the CLIP ids of the decoded caption are the concatenation of the CLIP ids of
its words (CLIP's pre-tokenizer splits on whitespace and punctuation, so neighbours do not interact).  The
exception is a '##' word-piece, which merges into the previous word; such ids are reported in
`needs_host` and captions containing them must take the host string path.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def build_bert2clip(bert_tokenizer, clip_tokenizer, vocab_size: int, special_ids) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    """Returns (off int32[V+1], tok int32[n], needs_host list of BERT ids that are '##' pieces)."""
    off = [0]
    toks: List[int] = []
    needs_host: List[int] = []
    special = set(int(s) for s in special_ids)
    conv = getattr(bert_tokenizer, "convert_ids_to_tokens", None)
    for v in range(vocab_size):
        if v not in special:
            piece = conv([v])[0] if conv else None
            if piece is not None and piece.startswith("##"):
                needs_host.append(v)
            text = bert_tokenizer.decode([v])
            if hasattr(clip_tokenizer, "tokens_of_text"):
                ids = clip_tokenizer.tokens_of_text(text)
            else:
                ids = clip_tokenizer(text, add_special_tokens=False)["input_ids"]
            toks.extend(int(i) for i in ids)
        off.append(len(toks))
    return torch.tensor(off, dtype=torch.int32), torch.tensor(toks, dtype=torch.int32), needs_host
