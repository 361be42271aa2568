"""CPU-side checks: the C-ABI library loads and exports every symbol include/conzic.h declares, refuses to
compute without a GPU (no fallback), and the host-side helpers behave like the reference's."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from conzic_b200 import _lib, tokens, utils
from synthetic import synth


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "conzic.h")).read()
    declared = sorted(set(re.findall(r"\b(conzic_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"libconzic.so does not export {name}"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.conzic_abi_version() == 6


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    cfg = _lib.Config()
    arr = (C.c_void_p * 10)()
    ctx = C.c_void_p()
    rc = lib.conzic_ctx_create(C.byref(cfg), arr, 10, arr, 5, None, C.byref(ctx))
    assert rc != 0 and "no CPU fallback" in _lib.last_error()
    from conzic_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine({}, {}, device="cuda:0")


def test_config_struct_matches_header_field_count():
    hdr = open(os.path.join(ROOT, "include", "conzic.h")).read()
    body = hdr.split("typedef struct conzic_config {")[1].split("} conzic_config;")[0]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    n = sum(len(d.split(",")) for d in re.findall(r"(?:int32_t|float)\s+([^;]+);", body))
    assert n == len(_lib.Config._fields_)
    body = hdr.split("typedef struct conzic_step_args {")[1].split("} conzic_step_args;")[0]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    n = sum(len(d.split(",")) for d in re.findall(r"(?:const\s+)?(?:int32_t|int64_t|float)\s*\*?\s*([^;]+);", body))
    assert n == len(_lib.StepArgs._fields_)


def test_init_text_and_dot_rule():
    tok = synth.SynthBertTokenizer()
    batch = utils.get_init_text(tok, synth.SYNTH_PROMPT, 4, batch_size=3)
    assert batch == [[101, 3746, 1997, 1037, 103, 103, 103, 103, 102]] * 3
    m = synth.make_token_mask()
    for i in range(4):
        out = utils.update_token_mask(tok, m, 4, i)
        assert out is m and float(m[0, synth.DOT_ID]) == (1.0 if i == 3 else 0.0)


def test_set_seed_reproduces_reference_visiting_orders():
    utils.set_seed(42)
    a = list(range(10)); random.shuffle(a)
    r = [np.random.randint(0, 10) for _ in range(5)]
    random.seed(42); np.random.seed(42)
    b = list(range(10)); random.shuffle(b)
    assert a == b and r == [np.random.randint(0, 10) for _ in range(5)]


@pytest.mark.parametrize("multi", [False, True])
def test_generic_table_builder_matches_synthetic_table(multi):
    """tokens.build_bert2clip walks the vocabulary through decode + CLIP tokenise; on the synthetic vocabularies it
    must reproduce the closed-form table."""
    V = 3000
    off, tok, needs_host = tokens.build_bert2clip(synth.SynthBertTokenizer(), synth.SynthCLIPTokenizer(multi), V,
                                                  synth.SPECIAL_IDS)
    off2, tok2 = synth.build_bert2clip_table(multi, vocab=V)
    assert torch.equal(off, off2) and torch.equal(tok, tok2) and needs_host == []
    for s in synth.SPECIAL_IDS:
        assert int(off[s + 1] - off[s]) == 0


def test_reference_api_surface():
    """Same public names and keyword defaults as gen_utils.py:289-292 / control_gen_utils.py:197-200."""
    import inspect
    from conzic_b200 import control_gen_utils, gen_utils
    sig = inspect.signature(gen_utils.generate_caption)
    assert list(sig.parameters)[:7] == ["img_name", "model", "clip", "tokenizer", "image_instance", "token_mask", "logger"]
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
    assert d == dict(prompt="", batch_size=1, max_len=15, top_k=100, temperature=1.0, max_iter=500, alpha=0.7, beta=1,
                     generate_order="sequential")
    sig = inspect.signature(control_gen_utils.control_generate_caption)
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
    assert d["gamma"] == 5 and d["ctl_type"] == "sentiment" and d["style_type"] == "positive" and d["max_len"] == 25
    sig = inspect.signature(gen_utils.sequential_generation)
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
    assert d == dict(max_len=15, top_k=100, temperature=None, alpha=0.7, beta=1, max_iters=20, batch_size=1, verbose=True)
    for name in ("shuffle_generation", "random_generation", "generate_caption_step"):
        assert hasattr(gen_utils, name)
    for name in ("sentiment_sequential_generation", "sentiment_shuffle_generation", "generate_caption_step",
                 "POS_sequential_generation"):
        assert hasattr(control_gen_utils, name)
    sig = inspect.signature(control_gen_utils.POS_sequential_generation)  # control_gen_utils.py:136-139
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
    assert d == dict(max_len=15, top_k=0, temperature=None, alpha=0.7, beta=1, gamma=0.1, max_iters=20, batch_size=1,
                     ctl_signal=["DET"], verbose=True)


def test_pos_template_scoring_matches_oracle_and_known_answers():
    """POS_classifier.py:6-31 through the product's host scorer: hand-worked cases (cut, pad, empty slot, list
    membership versus substring slots) and agreement with the oracle's restatement on synthetic captions."""
    from conzic_b200 import control_gen_utils as cgu
    from oracle import conzic_oracle as orc
    canned = {"a": ["DET", "NOUN"], "b": ["DET", "NOUN", "VERB", "ADV"], "c": ["NOUN"], "d": []}
    tagger = canned.__getitem__
    tpl = [["DET"], "", ["VERB", "NOUN"]]
    tags, sc = cgu.batch_texts_POS_analysis(["a", "b", "c", "d"], tpl, tagger=tagger)
    assert tags == [canned[k] for k in "abcd"]
    # a: DET ok, "" ok, pad "" not in list -> 2/3;  b: cut to 3, all ok -> 1;  c: NOUN not DET, "" ok, pad -> 1/3;  d: 1/3
    np.testing.assert_array_equal(sc.numpy(), np.array([2 / 3, 1.0, 1 / 3, 1 / 3], dtype=np.float32))
    # a string slot is a substring test, so the "" padding matches it (the reference's `in` quirk)
    _, sc = cgu.batch_texts_POS_analysis(["c", "d"], ["NOUN", "ADV"], tagger=tagger)
    np.testing.assert_array_equal(sc.numpy(), np.array([1.0, 1.0], dtype=np.float32))
    texts = [" ".join(f"w{2000 + (i * 37 + j * 101) % 20000}" for j in range(3 + i % 6)) + (" ." if i % 3 == 0 else "")
             for i in range(64)]
    t1, s1 = cgu.batch_texts_POS_analysis(texts, synth.SYNTH_POS_TEMPLATE, tagger=synth.synth_pos_tagger)
    t2, s2 = orc.pos_template_scores(texts, synth.SYNTH_POS_TEMPLATE, synth.synth_pos_tagger)
    assert t1 == t2 and torch.equal(s1, s2) and len(set(s1.tolist())) > 2


def test_pos_control_without_a_tagger_fails_loudly():
    from conzic_b200 import control_gen_utils as cgu
    cgu.set_pos_tagger(None)
    with pytest.raises(RuntimeError, match="set_pos_tagger"):
        cgu.batch_texts_POS_analysis(["w2000 w2001"], [["DET"]])


def test_cli_flags_and_defaults_match_the_reference():
    """run.py:15-76 / demo.py:15-76 of the reference: same flag names and defaults (demo differs in three)."""
    from conzic_b200 import cli
    r = cli.get_args("run", [])
    d = cli.get_args("demo", [])
    expect = dict(seed=42, device="cuda", run_type="controllable", prompt="Image of a", order="shuffle",
                  control_type="sentiment", sentiment_type="positive", samples_num=2, sentence_len=10, candidate_k=200,
                  alpha=0.02, beta=2.0, gamma=5.0, lm_temperature=0.1, num_iterations=10, lm_model="bert-base-uncased",
                  stop_words_path="stop_words.txt", add_extra_stopwords=[])
    for k, v in expect.items():
        assert getattr(r, k) == v and getattr(d, k) == v, k
    assert (r.batch_size, d.batch_size) == (2, 1)
    assert (r.match_model, d.match_model) == ("clip-vit-base-patch32", "openai/clip-vit-base-patch32")
    assert (r.caption_img_path, d.caption_img_path) == ("./examples/", "./examples/girl.jpg")
    assert len(r.pos_type) == 12 and r.pos_type[1] == ["ADJ", "NOUN"]
    a = cli.get_args("run", ["--run_type", "caption", "--order", "span", "--sentence_len", "12"])
    assert (a.run_type, a.order, a.sentence_len) == ("caption", "span", 12)


def test_root_level_modules_mirror_the_reference_layout():
    import importlib
    for name, attr in (("gen_utils", "generate_caption"), ("control_gen_utils", "control_generate_caption"),
                       ("utils", "set_seed"), ("clip.clip", "CLIP")):
        assert hasattr(importlib.import_module(name), attr)


def test_piece_vocabulary_decodes_like_hf_bert():
    """'##' pieces join the previous word (HF convert_tokens_to_string); a leading piece stays literal; the CLIP side
    tokenises a merged word as a whole, not as the concatenation of its parts."""
    tok, ctok = synth.PieceBertTokenizer(), synth.PieceCLIPTokenizer()
    assert synth.is_piece(2003) and not synth.is_piece(2004) and not synth.is_piece(synth.DOT_ID)
    ids = [101, 3746, 2003, 2008, 2004, 103, 1012, 102]
    assert tok.decode(ids) == "[CLS] w3746p2003p2008 w2004 [MASK] . [SEP]"
    assert tok.decode(ids, skip_special_tokens=True) == "w3746p2003p2008 w2004 ."
    assert tok.decode([2003, 2004], skip_special_tokens=True) == "##p2003 w2004"
    assert tok.vocab["##p2003"] == 2003 and tok.convert_ids_to_tokens([2003, 2004]) == ["##p2003", "w2004"]
    merged = ctok.tokens_of_text("w3746p2003")
    assert merged != ctok.tokens_of_text("w3746") + ctok.tokens_of_text("w2003") and len(merged) == 1


def test_sentiment_table_export_follows_the_reference_scoring():
    """tools/export_sentiment_table.py with stand-in NLTK callables: Penn tag -> WordNet class map and the mean of
    pos - neg over the synsets (sentiments_classifer.py:14-30); pieces and special tokens score 0."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("export_sentiment_table", os.path.join(ROOT, "tools", "export_sentiment_table.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tags = {"good": "JJ", "dog": "NN", "run": "VB", "the": "DT", "well": "RB"}
    syn = lambda p, n: types.SimpleNamespace(pos_score=lambda: p, neg_score=lambda: n)
    synsets = {("good", "a"): [syn(0.75, 0.0), syn(0.5, 0.25)], ("dog", "n"): [syn(0.0, 0.125)], ("run", "v"): [],
               ("well", "r"): [syn(0.5, 0.0)], ("the", ""): [syn(0.9, 0.0)]}
    calls = []

    def senti_synsets(w, t):
        calls.append((w, t))
        return synsets.get((w, t), [])

    tokens = ["[PAD]", "[unused5]", "good", "dog", "run", "the", "well", "##ing", "[MASK]"]
    table = mod.build_table(tokens, lambda ws: [(w, tags[w]) for w in ws], senti_synsets)
    np.testing.assert_allclose(table.numpy(), [0, 0, (0.75 + 0.25) / 2, -0.125, 0, 0.9, 0.5, 0, 0], atol=1e-7)
    assert ("good", "a") in calls and ("the", "") in calls and all(w not in ("##ing", "[PAD]") for w, _ in calls)


def test_table_with_real_hf_tokenizer_classes(tmp_path):
    """tokens.build_bert2clip with the REAL Hugging Face classes (BertTokenizer / CLIPTokenizer built from in-memory
    vocabularies; no pretrained files exist offline): whole-word tokens get the CLIP ids of their decoded text,
    '##' pieces get empty rows and are reported; the pair is recognised by the device text pipeline
    (tests/test_text_pipeline.py checks that pipeline against these classes caption by caption)."""
    import json
    from transformers import BertTokenizer, CLIPTokenizer
    base = ["[PAD]"] + [f"[unused{i}]" for i in range(1, 100)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    words = ["a", "dog", "cat", "run", "the", "girl", "red", "car", "tree", "walk", "small", "house", ".", "image", "of"]
    pieces = ["##s", "##ning", "##ed", "##er"]
    vocab = base + words + pieces
    (tmp_path / "vocab.txt").write_text("\n".join(vocab) + "\n")
    bt = BertTokenizer(str(tmp_path / "vocab.txt"))
    alphabet = list("abcdefghijklmnopqrstuvwxyz.")
    cv = {}
    for c in alphabet:
        cv[c] = len(cv)
    for c in alphabet:
        cv[c + "</w>"] = len(cv)
    merges = ["d o", "do g</w>", "c a", "ca t</w>", "t h", "th e</w>", "do g", "dog s</w>", "r e", "re d</w>", "e d</w>",
              "c a", "ca r</w>", "ca r", "car s</w>", "e r</w>", "w a", "wa l", "wal k</w>", "wal k", "walk ed</w>"]
    merges = list(dict.fromkeys(merges))
    for m in merges:
        cv[m.replace(" ", "")] = len(cv)
    cv["<|startoftext|>"] = len(cv)
    cv["<|endoftext|>"] = len(cv)
    (tmp_path / "cv.json").write_text(json.dumps(cv))
    (tmp_path / "merges.txt").write_text("#version: 0.2\n" + "\n".join(merges) + "\n")
    ct = CLIPTokenizer(str(tmp_path / "cv.json"), str(tmp_path / "merges.txt"), model_max_length=77)
    V = len(vocab)
    special = [0, 100, 101, 102, 103]
    off, tk, piece_ids = tokens.build_bert2clip(bt, ct, V, special)
    assert [vocab[v] for v in piece_ids] == pieces
    table = lambda v: tk[off[v]: off[v + 1]].tolist()
    for w in words:
        assert table(vocab.index(w)) == ct(w, add_special_tokens=False)["input_ids"]
    for v in piece_ids + special:
        assert table(v) == []
    ok, why = tokens.text_vocab_supported(bt, ct)
    assert ok, why
    tv = tokens.build_text_vocab(bt, ct, V, special, off, tk)
    fl = tv["tok_flags"]
    assert all(int(fl[vocab.index(p)]) & 1 for p in pieces) and int(fl[vocab.index(".")]) & 2
    assert int(fl[vocab.index("dog")]) & 8 and not int(fl[vocab.index("##s")]) & 8 and int(fl[101]) & 4
    ok, why = tokens.text_vocab_supported(synth.SynthBertTokenizer(), synth.SynthCLIPTokenizer())
    assert not ok and "fast tokenizers" in why
