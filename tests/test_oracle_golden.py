"""The CPU oracle replayed against fixtures recorded from the UNMODIFIED reference
(tools/make_golden.py).  This is what pins oracle/conzic_oracle.py."""
import os
import random
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from synthetic import synth
from oracle import conzic_oracle as orc

ALL = sorted(f[:-3] for f in os.listdir(GOLDEN) if f.endswith(".pt"))
FAST = ["seq_b2_n4_k8", "shuffle_b3_n5_k16_multi", "senti_shuffle_neg_b2_n4_k8", "peaked_seq_b2_n4_k32",
        "pos_seq_b2_n5_k16", "pieces_shuffle_b3_n6_k16_multi"]


def _pos(case):
    if case.get("ctl") != "pos":
        return {}
    return dict(pos_template=synth.SYNTH_POS_TEMPLATE, pos_tagger=synth.synth_pos_tagger)


def make_oracle(g, synth_weights, full_logits=False):
    case = g["case"]
    bert_sd = synth_weights("bert", case.get("peaked", False))
    clip_sd = synth_weights("clip")
    assert synth.state_dict_checksum(bert_sd) == g["bert_crc"], "synthetic BERT weights drifted"
    assert synth.state_dict_checksum(clip_sd) == g["clip_crc"], "synthetic CLIP weights drifted"
    multi = case.get("multi", False)
    if case.get("hf"):  # the real transformers tokenizer classes over generated vocabulary files
        import tempfile
        toks = synth.make_hf_tokenizers(tempfile.mkdtemp())
    elif case.get("pieces"):  # vocabulary with '##' word pieces
        toks = synth.PieceBertTokenizer(), synth.PieceCLIPTokenizer(multi)
    else:
        toks = synth.SynthBertTokenizer(), synth.SynthCLIPTokenizer(multi)
    return orc.Oracle(bert_sd, clip_sd, toks[0], toks[1], sentiment_table=synth.make_sentiment_table(),
                      full_logits=full_logits)


@pytest.mark.parametrize("name", FAST)
def test_teacher_forced_steps(name, synth_weights):
    """Every recorded step, fed the reference's own ``inp``: logits, top-k, CLIP ids, embeds,
    similarity, and the winner (visible in the next step's ``inp``) must agree."""
    g = load_golden(name)
    case = g["case"]
    o = make_oracle(g, synth_weights)
    o.trace = []
    n, K = case["n"], case["K"]
    steps = g["steps"]
    torch.manual_seed(0)
    for si, s in enumerate(steps):
        inp = s["inp"].clone()
        pos = s["pos"]
        ii = pos - 4
        token_mask = synth.make_token_mask()
        with torch.no_grad():
            o.step(inp, s["image_embeds"], token_mask, pos, ii, n, K, 0.1, 0.02, 2.0,
                   gamma=case.get("gamma"), ctl_signal=case.get("style", "positive"), **_pos(case))
        t = o.trace[-1]
        assert float(token_mask[0, synth.DOT_ID]) == s["token_mask_dot"]
        cols = s["logit_cols"].long()
        torch.testing.assert_close(t["logits_row"].gather(1, cols), s["logit_vals"], rtol=0, atol=2e-5)
        # top-k: identical ids wherever the reference's probabilities are distinct and non-zero
        torch.testing.assert_close(t["probs"], s["probs"], rtol=1e-4, atol=1e-30)
        nz = s["probs"] > 0
        assert torch.equal(t["idxs"][nz], s["idxs"][nz])
        assert torch.equal(t["clip_ids"][:, : s["clip_ids"].shape[1]].int(), s["clip_ids"]) or not bool(nz.all())
        if bool(nz.all()):
            torch.testing.assert_close(t["text_embeds"][s["embed_rows"]], s["text_embeds"], rtol=0, atol=2e-5)
            torch.testing.assert_close(t["clip_ref"], s["clip_ref"], rtol=0, atol=2e-6)
            torch.testing.assert_close(t["clip_score"], s["clip_score"], rtol=1e-4, atol=1e-7)
            if si + 1 < len(steps):
                nxt = steps[si + 1]["inp"]
                keep = torch.ones(inp.shape[1], dtype=torch.bool)
                keep[steps[si + 1]["pos"]] = False  # the next step masks its own position
                assert torch.equal(inp[:, keep], nxt[:, keep])


@pytest.mark.parametrize("name", ["seq_b2_n4_k8", "random_b2_n3_k8", "senti_seq_b2_n4_k8", "span_b2_n5_k8",
                                  "pos_seq_b2_n5_k16", "pieces_seq_b2_n5_k16", "hf_shuffle_b3_n5_k24"])
def test_free_running_call(name, synth_weights):
    """Whole ``generate_caption`` / ``control_generate_caption`` call under set_seed(42):
    same captions per sweep, same CLIP scores, same best list as the reference returned."""
    g = load_golden(name)
    case = g["case"]
    o = make_oracle(g, synth_weights, full_logits=(name == "seq_b2_n4_k8"))
    random.seed(42); np.random.seed(42); torch.manual_seed(42)  # utils.py:37-44
    pix = torch.stack([synth.make_pixel_values(i) for i in range(case["B"])])
    with torch.no_grad():
        prompt = synth.hf_prompt() if case.get("hf") else synth.SYNTH_PROMPT
        texts, scores = o.generate(pix, synth.make_token_mask(), prompt, order=case["order"],
                                   max_len=case["n"], top_k=case["K"], max_iters=case["iters"],
                                   gamma=case.get("gamma"), ctl_signal=case.get("style", "positive"), **_pos(case))
    assert texts == g["texts"]
    assert len(scores) == len(g["scores"])
    for a, b in zip(scores, g["scores"]):
        np.testing.assert_allclose(a, b, rtol=0, atol=3e-6)


def test_image_embeds_match_reference(synth_weights):
    g = load_golden("seq_b2_n4_k8")
    pix = torch.stack([synth.make_pixel_values(i) for i in range(2)])
    with torch.no_grad():
        e = orc.clip_image_embeds(synth_weights("clip"), pix)
    torch.testing.assert_close(e, g["steps"][0]["image_embeds"], rtol=0, atol=2e-5)


def test_fixture_inventory():
    assert set(FAST) <= set(ALL) and "seq_b1_n10_k200" in ALL
