"""BERT-id -> CLIP-BPE-id table: the device-side replacement of the reference's per-step string round
trip `tokenizer.batch_decode(...)` -> `CLIPTokenizer(...)` (gen_utils.py:75, clip/clip.py:71-72).

For a word-level BERT token the CLIP ids of the decoded caption are the concatenation of the CLIP ids of
its words (CLIP's pre-tokenizer splits on whitespace and punctuation, so neighbours do not interact).  The
exception is a '##' word-piece, which merges into the previous word and changes that word's BPE: such ids get an
EMPTY table row and are reported as pieces.  Vocabularies that have them (every real BERT vocabulary) run through
the device text pipeline (csrc/text_pipeline.cuh; `build_text_vocab` below builds its tables): WordPiece decode,
CLIP pre-tokenisation and byte-level BPE per candidate caption on the device.
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


# ---------------------------------------------------------------------------------------------------------------
# Device text pipeline (csrc/text_pipeline.cuh): byte tables of the BERT vocabulary + the CLIP BPE model
# ---------------------------------------------------------------------------------------------------------------
CLIP_SPLIT_REGEX = r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"
_GLUE_PREFIXES = (".", "?", "!", ",", "n't", "'m", "'s", "'ve", "'re")  # tokenizers decoders/wordpiece.rs cleanup()


def bytes_to_unicode():
    """The GPT-2 / CLIP ByteLevel table: every byte as a printable unicode character (tokenizers pre_tokenizers/byte_level.rs)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


def _char_class(ch: str) -> int:
    import unicodedata
    if ch.isspace():
        return 3
    cat = unicodedata.category(ch)
    return 1 if cat[0] == "L" else (2 if cat[0] == "N" else 0)


def text_vocab_supported(bert_tokenizer, clip_tokenizer):
    """(ok, why): the device text pipeline restates exactly one tokenizer pair -- the Hugging Face fast
    BertTokenizer (WordPiece decoder, prefix '##', per-token clean-up, no whole-string clean-up on top) and
    CLIPTokenizer (NFC / whitespace / lowercase, the CLIP split regex, ByteLevel, BPE with '</w>').  Anything else
    (duck-typed tokenizers, other decoders) keeps the table path, or the string path when it has '##' pieces."""
    import json
    bk, ck = getattr(bert_tokenizer, "backend_tokenizer", None), getattr(clip_tokenizer, "backend_tokenizer", None)
    if bk is None or ck is None:
        return False, "not Hugging Face fast tokenizers"
    try:
        dec = json.loads(bk.decoder.__getstate__())
    except Exception:  # noqa: BLE001
        return False, "BERT tokenizer has no inspectable decoder"
    if dec.get("type") != "WordPiece" or dec.get("prefix") != "##" or not dec.get("cleanup", False):
        return False, f"unsupported BERT decoder {dec}"
    if getattr(bert_tokenizer, "clean_up_tokenization_spaces", False):
        return False, "BERT tokenizer applies the whole-string clean_up_tokenization_spaces pass"
    cj = json.loads(ck.to_str())
    m = cj.get("model", {})
    if (m.get("type") != "BPE" or m.get("end_of_word_suffix") != "</w>" or m.get("continuing_subword_prefix") not in ("", None)
            or m.get("byte_fallback") or m.get("dropout") or m.get("ignore_merges")):
        return False, "unsupported CLIP tokenizer model"
    norm = [n.get("type") for n in (cj.get("normalizer") or {}).get("normalizers", [])]
    pre = (cj.get("pre_tokenizer") or {}).get("pretokenizers", [])
    if norm != ["NFC", "Replace", "Lowercase"] or len(pre) != 2 or pre[0].get("type") != "Split" \
            or pre[0].get("pattern", {}).get("Regex") != CLIP_SPLIT_REGEX or pre[1].get("type") != "ByteLevel" \
            or pre[1].get("add_prefix_space"):
        return False, "unsupported CLIP normaliser / pre-tokeniser"
    return True, ""


def build_text_vocab(bert_tokenizer, clip_tokenizer, vocab_size: int, special_ids, csr_off: torch.Tensor,
                     csr_tok: torch.Tensor):
    """Host arrays behind conzic_set_text_vocab: per-token byte / character-class tables of the BERT vocabulary
    (normalised the way the CLIP tokenizer normalises text: NFC, lowercase) and the CLIP BPE model (byte symbols,
    merges as an open-addressing hash table).  `csr_off`, `csr_tok`: the per-token CLIP ids from build_bert2clip,
    used as a shortcut for "simple" tokens (their text alone is exactly one pre-token).  Returns a dict of tensors."""
    import json
    import unicodedata
    special = set(int(s) for s in special_ids) | set(int(s) for s in getattr(bert_tokenizer, "all_special_ids", []))
    toks = bert_tokenizer.convert_ids_to_tokens(list(range(vocab_size)))
    off, data, cls, flags = [0], bytearray(), bytearray(), bytearray(vocab_size)
    for v, t in enumerate(toks):
        fl = 0
        if t is None or v in special:
            fl |= 4
            t = ""
        if t.startswith("##") and len(t) > 2:
            fl |= 1
            t = t[2:]
        if any(t.startswith(p) for p in _GLUE_PREFIXES):
            fl |= 2
        t = unicodedata.normalize("NFC", t).lower()
        classes = [_char_class(ch) for ch in t]
        # simple: the whole text is one pre-token of the CLIP regex (a run of letters, one number character, or a run of
        # other characters that does not start with a contraction) -- then its own CLIP ids can be copied
        simple = bool(t) and not (fl & 5)  # pieces have no row in the per-token table
        if simple:
            if all(c == 1 for c in classes):
                pass
            elif len(t) == 1 and classes[0] == 2:
                pass
            elif all(c == 0 for c in classes):
                simple = not (t[0] == "'" and len(t) > 1)  # other runs hold no letters, so no contraction can start inside
            else:
                simple = False
        if simple:
            fl |= 8
        for ch, c in zip(t, classes):
            b = ch.encode("utf-8")
            data += b
            cls += bytes([c | 4] + [c] * (len(b) - 1))
        off.append(len(data))
        flags[v] = fl
    cj = json.loads(clip_tokenizer.backend_tokenizer.to_str())
    model = cj["model"]
    vocab = model["vocab"]
    unk = vocab.get(model.get("unk_token") or "", 0)
    b2u = bytes_to_unicode()
    byte_sym = [vocab.get(b2u[b], unk) for b in range(256)] + [vocab.get(b2u[b] + "</w>", unk) for b in range(256)]
    merges = [m.split(" ") if isinstance(m, str) else list(m) for m in model["merges"]]
    if len(merges) >= 65535 or max(vocab.values()) >= 65536:
        raise ValueError("CLIP BPE model too large for the packed merge table (rank / id must fit 16 bits)")
    bits = max(4, (2 * max(len(merges), 1) - 1).bit_length())
    size = 1 << bits
    keys, vals = [0] * size, [0] * size
    mask64 = (1 << 64) - 1
    for rank, (a, b) in enumerate(merges):
        ia, ib, iab = vocab.get(a), vocab.get(b), vocab.get(a + b)
        if ia is None or ib is None or iab is None:
            continue  # the library rejects such files; generated test vocabularies are consistent
        key = ((ia << 32) | ib) + 1
        slot = ((key * 0x9E3779B97F4A7C15) & mask64) >> (64 - bits)
        while keys[slot] != 0:
            if keys[slot] == key:
                break
            slot = (slot + 1) & (size - 1)
        if keys[slot] == key:
            continue  # an earlier (lower-rank) merge of the same pair wins
        keys[slot], vals[slot] = key, (rank << 16) | iab
    import numpy as np
    return dict(tok_off=torch.tensor(off, dtype=torch.int32), tok_bytes=torch.tensor(list(data) or [0], dtype=torch.uint8),
                tok_cls=torch.tensor(list(cls) or [0], dtype=torch.uint8),
                tok_flags=torch.tensor(list(flags), dtype=torch.uint8),
                csr_off=csr_off.to(torch.int32), csr_tok=csr_tok.to(torch.int32) if csr_tok.numel() else torch.zeros(1, dtype=torch.int32),
                byte_sym=torch.tensor(byte_sym, dtype=torch.int32),
                merge_keys=torch.from_numpy(np.array(keys, dtype=np.uint64).view(np.int64)),
                merge_vals=torch.from_numpy(np.array(vals, dtype=np.uint32).view(np.int32)), merge_bits=bits)
