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
    needs_host: List[int] = []
    special = set(int(s) for s in special_ids)
    conv = getattr(bert_tokenizer, "convert_ids_to_tokens", None)
    pieces = conv(list(range(vocab_size))) if conv else [None] * vocab_size
    texts: List[str] = []
    owners: List[int] = []
    for v in range(vocab_size):
        if v in special:
            continue
        if pieces[v] is not None and pieces[v].startswith("##"):
            needs_host.append(v)
            continue
        texts.append(bert_tokenizer.decode([v]))
        owners.append(v)
    if hasattr(clip_tokenizer, "tokens_of_text"):
        rows = [clip_tokenizer.tokens_of_text(t) for t in texts]
    else:  # one batched call: HF tokenizers take a list (30 k single calls cost tens of seconds)
        rows = clip_tokenizer(texts, add_special_tokens=False)["input_ids"] if texts else []
    per_id = [()] * vocab_size
    for v, r in zip(owners, rows):
        per_id[v] = r
    off = [0]
    toks: List[int] = []
    for v in range(vocab_size):
        toks.extend(int(i) for i in per_id[v])
        off.append(len(toks))
    return torch.tensor(off, dtype=torch.int32), torch.tensor(toks, dtype=torch.int32), needs_host


def hybrid_flags(inp_h: torch.Tensor, pos: int, ids_masked_h: torch.Tensor, piece: torch.Tensor, special_ids):
    """Which candidate captions need the host string pass, and which images only need their prefix / tail re-tokenised.

    A caption whose words outside `pos` hold a piece is still the space-joined sequence prefix + candidate + tail as
    long as no piece follows `pos` directly (that one would merge INTO the candidate word, also across a dropped
    candidate): such an image keeps the table path for its candidates once the host supplies the CLIP ids of its
    prefix and tail strings.  Returns (flag bool[B,K], override bool[B]):
      flag[b,k]   -- candidate k is a piece itself, or image b cannot be overridden (a piece right after `pos`);
      override[b] -- image b holds a piece outside `pos` and can be handled by a prefix / tail override."""
    others = inp_h.clone()
    others[:, pos] = int(special_ids[0])
    pm = piece[others]
    img_flag = pm.any(dim=1)
    kept = ~torch.isin(others, torch.as_tensor(list(special_ids), dtype=others.dtype))
    after = kept[:, pos + 1:]
    if after.shape[1] > 0:
        first = after.int().argmax(dim=1)
        next_piece = after.any(dim=1) & pm[:, pos + 1:].gather(1, first.view(-1, 1)).squeeze(1)
    else:
        next_piece = torch.zeros_like(img_flag)
    override = img_flag & ~next_piece
    flag = piece[ids_masked_h] | (img_flag & ~override)[:, None]
    return flag, override


def hybrid_capacities(inp_h: torch.Tensor, pos: int, ids_masked_h: torch.Tensor, tok_len: torch.Tensor, ov_lens=None,
                      maxpos: int = 77):
    """Row capacities (P incl. BOS, S incl. EOS) for conzic_encode_candidates; `ov_lens` maps an overridden image
    index to the (prefix, tail) token counts of its host-tokenised strings."""
    others = inp_h.clone()
    others[:, pos] = 0
    lens = tok_len[others].long()
    pre = 1 + lens[:, :pos].sum(dim=1)
    tail = lens[:, pos + 1:].sum(dim=1)
    for b, (n_pre, n_tail) in (ov_lens or {}).items():
        pre[b], tail[b] = 1 + n_pre, n_tail
    cand = tok_len[ids_masked_h].long()
    cap = maxpos - 1
    P = int(min(max(int(pre.max()), 1), cap))
    S = int(min(max(int((cand + tail[:, None]).max()) + 1, 2), cap))
    return P, S
