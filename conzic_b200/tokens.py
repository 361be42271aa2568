"""BERT-id -> CLIP-BPE-id table: the device-side replacement of the reference's per-step string round
trip `tokenizer.batch_decode(...)` -> `CLIPTokenizer(...)` (gen_utils.py:75, clip/clip.py:71-72).

For a word-level BERT token the CLIP ids of the decoded caption are the concatenation of the CLIP ids of
its words (CLIP's pre-tokenizer splits on whitespace and punctuation, so neighbours do not interact).  The
exception is a '##' word-piece, which merges into the previous word and changes that word's BPE: such ids get an
EMPTY table row and are reported in `needs_host`.  Captions that contain one are re-encoded from their strings by
the host and patched over the table result (`plan_hybrid` decides which).
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
                off.append(len(toks))
                continue
            text = bert_tokenizer.decode([v])
            if hasattr(clip_tokenizer, "tokens_of_text"):
                ids = clip_tokenizer.tokens_of_text(text)
            else:
                ids = clip_tokenizer(text, add_special_tokens=False)["input_ids"]
            toks.extend(int(i) for i in ids)
        off.append(len(toks))
    return torch.tensor(off, dtype=torch.int32), torch.tensor(toks, dtype=torch.int32), needs_host


def plan_hybrid(inp_h: torch.Tensor, pos: int, ids_masked_h: torch.Tensor, piece: torch.Tensor, tok_len: torch.Tensor,
                maxpos: int = 77):
    """Host-side plan of one step for a vocabulary with '##' pieces.

    inp_h int64[B,L] (column `pos` is ignored), ids_masked_h int64[B,K] (candidate ids, 0 where masked),
    piece bool[V], tok_len int[V] (CLIP tokens per BERT id; 0 for special ids and pieces).
    Returns (flag bool[B,K], P, S): flag marks the candidate captions that contain a piece anywhere -- the table
    result is not valid for them; P / S are the exact row capacities of the shared prefix (BOS included) and of
    the per-candidate part (EOS included) in CLIP tokens, clamped to what a 77-token caption can hold."""
    others = inp_h.clone()
    others[:, pos] = 0  # [PAD]: special, no piece, no tokens
    img_flag = piece[others].any(dim=1)
    flag = piece[ids_masked_h] | img_flag[:, None]
    lens = tok_len[others].long()
    pre = 1 + lens[:, :pos].sum(dim=1)
    tail = lens[:, pos + 1:].sum(dim=1)
    cand = tok_len[ids_masked_h].long()
    cap = maxpos - 1
    P = int(min(max(int(pre.max()), 1), cap))
    S = int(min(max(int((cand + tail[:, None]).max()) + 1, 2), cap))
    return flag, P, S
