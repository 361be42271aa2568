"""Imports the UNMODIFIED reference modules (gen_utils, control_gen_utils, utils, clip/clip.py, POS_classifier)
from a directory -- /root/reference in the build container (tools/make_golden.py) or oracle/_ref on the GPU box
(bench.py --impl reference) -- with the shims SURVEY.md 8(c) lists for what is not installed offline: a one-line
`colorlog`, a table-driven `sentiments_classifer`, and the two NLTK callables POS_classifier.py uses.
Test infrastructure: only tools/, tests/ and bench.py's reference arm import this."""
import importlib
import importlib.util
import logging
import os
import sys
import types

import torch


def load(ref_dir, sentiment_table=None, word_to_id=None, pos_tagger=None):
    """Returns (utils, gen_utils, control_gen_utils, CLIP) of the reference at `ref_dir`.
    sentiment_table f32[V] + word_to_id(word) -> id back the stub sentiment scorer (same maths as
    sentiments_classifer.py:35-48 with a per-word table for SentiWordNet); pos_tagger(text) -> tags backs nltk.pos_tag."""
    ref_dir = os.path.abspath(ref_dir)
    sys.path.insert(0, ref_dir)
    for name in ("utils", "gen_utils", "control_gen_utils", "POS_classifier"):
        sys.modules.pop(name, None)  # this repo's root-level shims carry the same names
    sys.modules["colorlog"] = types.SimpleNamespace(
        ColoredFormatter=lambda *a, **k: logging.Formatter("%(message)s"))

    def table_scorer(batch_texts, temperature, device, sentiment_ctl=None, batch_size_image=1):
        s = torch.zeros(len(batch_texts))
        for i, t in enumerate(batch_texts):
            v = sum(float(sentiment_table[word_to_id(w)]) for w in t.split())
            s[i] = -v if sentiment_ctl == "negative" else v
        sb = s.view(batch_size_image, -1).to(device)
        return torch.softmax(sb / temperature, dim=1).to(device), sb, [], []

    sys.modules["sentiments_classifer"] = types.SimpleNamespace(batch_texts_POS_Sentiments_analysis=table_scorer)
    nltk = types.ModuleType("nltk")
    nltk.tokenize = types.ModuleType("nltk.tokenize")
    nltk.tokenize.word_tokenize = lambda text: text.replace(".", " . ").split()
    nltk.pos_tag = lambda words, tagset=None: list(zip(words, pos_tagger(" ".join(words))))
    sys.modules["nltk"], sys.modules["nltk.tokenize"] = nltk, nltk.tokenize
    utils = importlib.import_module("utils")
    gen_utils = importlib.import_module("gen_utils")
    control_gen_utils = importlib.import_module("control_gen_utils")
    # the reference's clip/ has no __init__.py, so this repo's root-level `clip` shim package would win the
    # name; load the reference's file by path instead
    spec = importlib.util.spec_from_file_location("reference_clip_clip", os.path.join(ref_dir, "clip", "clip.py"))
    ref_clip = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_clip)
    for m in (utils, gen_utils, control_gen_utils):
        assert os.path.abspath(m.__file__).startswith(ref_dir), m.__file__
    return utils, gen_utils, control_gen_utils, ref_clip.CLIP


def build_models(bert_sd, clip_sd, CLIP, clip_tokenizer, processor):
    """HF BertForMaskedLM / CLIPModel (class defaults = bert-base-uncased / ViT-B/32 shapes) loaded with the given
    state dicts, and the reference's CLIP wrapper built around the CLIPModel without from_pretrained."""
    from transformers import BertConfig, BertForMaskedLM, CLIPConfig, CLIPModel
    bert = BertForMaskedLM(BertConfig()).eval()
    missing = bert.load_state_dict(bert_sd, strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys), missing
    clipm = CLIPModel(CLIPConfig()).eval()
    missing = clipm.load_state_dict(clip_sd, strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys), missing
    clip = CLIP.__new__(CLIP)
    torch.nn.Module.__init__(clip)
    clip.model, clip.processor, clip.tokenizer = clipm, processor, clip_tokenizer
    clip.cuda_has_been_checked = False
    return bert, clip
