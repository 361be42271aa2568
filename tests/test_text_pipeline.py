"""csrc/text_pipeline.cuh (BERT ids -> WordPiece decode -> CLIP normalise / pre-tokenise / BPE, the code the CUDA
kernels compile) built for the CPU and checked against the REAL transformers tokenizer classes:
`CLIPTokenizer(BertTokenizer.batch_decode(ids, skip_special_tokens=True))` (gen_utils.py:75 + clip/clip.py:71-72)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from conzic_b200 import tokens
from synthetic import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "harness", "libtext_host_harness.so")


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(HERE, "harness", "text_host_harness.cpp")
    hdr = os.path.join(HERE, "..", "conzic_b200", "csrc", "text_pipeline.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", SO], check=True)
    lib = C.CDLL(SO)
    lib.conzic_text_host_tokenize.restype = C.c_int
    return lib


def _run(lib, tv, special, rows):
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    n, L = rows.shape
    out = np.zeros((n, 75), dtype=np.int32)
    lens = np.zeros(n, dtype=np.int32)
    sp = (C.c_int * 5)(*special)
    keep = {k: v.numpy() for k, v in tv.items() if isinstance(v, torch.Tensor)}
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    err = lib.conzic_text_host_tokenize(p(keep["tok_off"]), p(keep["tok_bytes"]), p(keep["tok_cls"]), p(keep["tok_flags"]),
                                        p(keep["csr_off"]), p(keep["csr_tok"]), p(keep["byte_sym"]), p(keep["merge_keys"]),
                                        p(keep["merge_vals"]), C.c_int(tv["merge_bits"]), C.c_int(keep["tok_flags"].shape[0]),
                                        p(rows), C.c_int(n), C.c_int(L), sp, p(out), p(lens))
    return err, [out[i, : lens[i]].tolist() for i in range(n)]


def _reference(bert_tok, clip_tok, rows):
    texts = bert_tok.batch_decode(torch.as_tensor(rows), skip_special_tokens=True)
    ids = clip_tok(texts, add_special_tokens=False)["input_ids"]
    return texts, [r[:75] for r in ids]


def _vocab(bert_tok, clip_tok, V):
    special = [synth.PAD_ID, synth.UNK_ID, synth.CLS_ID, synth.SEP_ID, synth.MASK_ID]
    ok, why = tokens.text_vocab_supported(bert_tok, clip_tok)
    assert ok, why
    off, tok, _ = tokens.build_bert2clip(bert_tok, clip_tok, V, special)
    return tokens.build_text_vocab(bert_tok, clip_tok, V, special, off, tok), special


def test_letter_coded_vocabulary_with_pieces(harness, tmp_path):
    """The vocabulary of the hf_* golden fixtures: words, '##' pieces, '.', specials and [unusedN]."""
    bert_tok, clip_tok = synth.make_hf_tokenizers(str(tmp_path))
    tv, special = _vocab(bert_tok, clip_tok, synth.BERT_VOCAB)
    g = np.random.default_rng(0)
    rows = g.integers(1996, synth.BERT_VOCAB, size=(3000, 16))
    rows[:, 0], rows[:, -1] = synth.CLS_ID, synth.SEP_ID
    rows[g.random(rows.shape) < 0.08] = synth.DOT_ID
    rows[g.random(rows.shape) < 0.05] = synth.MASK_ID
    rows[g.random(rows.shape) < 0.03] = synth.PAD_ID
    rows[g.random(rows.shape) < 0.01] = 7          # an [unused7]
    rows[:50, 1] = [1996 + 5 * i + 2 for i in range(50)]  # ids % 5 == 3: pieces right after [CLS] (kept verbatim)
    err, got = _run(harness, tv, special, rows)
    texts, ref = _reference(bert_tok, clip_tok, rows)
    assert err == 0
    bad = [i for i in range(len(ref)) if got[i] != ref[i]]
    assert not bad, (texts[bad[0]], got[bad[0]], ref[bad[0]])
    assert any("##" in t for t in texts[:50])


def test_rich_vocabulary_punctuation_contractions_digits_unicode(harness, tmp_path):
    """Punctuation runs ('. .' -> '..'), contractions across token boundaries (it ' s, don ##'t), digits, multi-byte
    characters, pieces of every kind -- against the real tokenizer classes over the full byte alphabet."""
    bert_tok, clip_tok = synth.make_hf_tokenizers_rich(str(tmp_path))
    V = 3000
    tv, special = _vocab(bert_tok, clip_tok, V)
    g = np.random.default_rng(1)
    n_words = 58
    rows = g.integers(1996, V, size=(6000, 20))
    dense = g.random(rows.shape) < 0.6  # favour the hand-written entries (first 2 x 58 ids after 1996)
    rows[dense] = g.integers(1996, 1996 + 2 * n_words, size=int(dense.sum()))
    rows[:, 0], rows[:, -1] = synth.CLS_ID, synth.SEP_ID
    rows[g.random(rows.shape) < 0.05] = synth.DOT_ID
    rows[g.random(rows.shape) < 0.04] = synth.PAD_ID
    rows[g.random(rows.shape) < 0.01] = 55
    err, got = _run(harness, tv, special, rows)
    texts, ref = _reference(bert_tok, clip_tok, rows)
    assert err == 0
    bad = [i for i in range(len(ref)) if got[i] != ref[i]]
    assert not bad, (len(bad), texts[bad[0]], bert_tok.convert_ids_to_tokens(rows[bad[0]].tolist()), got[bad[0]], ref[bad[0]])
    joined = " ".join(texts)
    assert "'s" in joined and ".." in joined and "é" in joined


def test_long_captions_truncate_like_the_tokenizer(harness, tmp_path):
    bert_tok, clip_tok = synth.make_hf_tokenizers(str(tmp_path))
    tv, special = _vocab(bert_tok, clip_tok, synth.BERT_VOCAB)
    g = np.random.default_rng(2)
    rows = g.integers(1996, synth.BERT_VOCAB, size=(200, 40))  # ~4 CLIP tokens per word: well past 75
    err, got = _run(harness, tv, special, rows)
    _, ref = _reference(bert_tok, clip_tok, rows)
    assert err == 0 and got == ref and max(len(r) for r in got) == 75
