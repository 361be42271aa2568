"""Shared helpers for the -m gpu parity tests and tools/gpu_diag.py: build engines on synthetic weights,
replay golden fixtures through the C ABI, measure error magnitudes against the CPU oracle."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from synthetic import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
_SD = {}
_ENG = {}


def weights(kind, peaked=False):
    key = (kind, peaked)
    if key not in _SD:
        _SD[key] = synth.make_bert_state_dict(0, peaked=peaked) if kind == "bert" else synth.make_clip_state_dict(0)
    return _SD[key]


def engine(precision="bf16x3", impl="tcgen05", peaked=False, multi=False, **kw):
    """Engines are cached per configuration (weights upload takes a few seconds).  `kw`: Engine keyword
    arguments (cert_dcos, cert_fcap, ln_standalone ...)."""
    from conzic_b200.engine import Engine
    key = (precision, impl, peaked, multi, tuple(sorted(kw.items())))
    if key not in _ENG:
        e = Engine(weights("bert", peaked), weights("clip"), device="cuda:0", precision=precision, gemm_impl=impl, **kw)
        off, tok = synth.build_bert2clip_table(multi)
        e.set_bert2clip(off, tok)
        _ENG[key] = e
    return _ENG[key]


def drop_engines():
    for e in _ENG.values():
        e.close()
    _ENG.clear()
    torch.cuda.empty_cache()


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def visible_counts(inp_row_batch: torch.Tensor, pos: int):
    """Upper bounds the host loop would pass: words before / after `pos` (special ids hold no word)."""
    special = torch.tensor(synth.SPECIAL_IDS)
    vis = ~torch.isin(inp_row_batch, special)
    before = int(vis[:, :pos].sum(dim=1).max())
    after = int(vis[:, pos + 1:].sum(dim=1).max())
    return before, after


def replay_fixture(eng, g, max_steps=None):
    """Teacher-forced replay of every recorded step of a golden fixture through conzic_gibbs_step.
    Returns a dict of worst-case error magnitudes and mismatch counts.  CLIP cosines are compared candidate by
    candidate wherever the same id is in both top-k lists (the cosine depends on the caption only); the softmax
    score and the winner of an image are compared whenever that image's top-k id SET matches the reference's."""
    case = g["case"]
    n, K = case["n"], case["K"]
    gamma = case.get("gamma")
    table = synth.make_sentiment_table()
    if case.get("style") == "negative":
        table = -table
    dev = eng.device
    table_d = table.to(dev) if gamma is not None else None
    m = dict(steps=0, logit_err=0.0, prob_relerr=0.0, topk_id_mismatch=0, clip_ref_err=0.0, clip_score_err=0.0,
             winner_cos_err=0.0, winner_mismatch=0, winner_checked=0, min_margin_at_mismatch=None, dot_mismatch=0,
             cos_checked=0)
    steps = g["steps"][:max_steps] if max_steps else g["steps"]
    for si, s in enumerate(steps):
        inp = s["inp"].clone().to(dev)
        pos = s["pos"]
        ii = pos - 4
        token_mask = synth.make_token_mask(dev)
        before, after = visible_counts(s["inp"], pos)
        img = s["image_embeds"].to(dev).contiguous()
        clip_ref, senti, tr = eng.gibbs_step(inp, token_mask, img, pos, ii == n - 1, K, 0.1, 0.02, 2.0, before, after,
                                             gamma=gamma, senti_table=table_d, trace=True)
        torch.cuda.synchronize()
        tr = {k: v.cpu() for k, v in tr.items()}
        clip_ref = clip_ref.cpu()
        m["steps"] += 1
        if float(token_mask[0, synth.DOT_ID]) != s["token_mask_dot"]:
            m["dot_mismatch"] += 1
        cols = s["logit_cols"].long()
        m["logit_err"] = max(m["logit_err"], float((tr["logits"][:, : eng.V].gather(1, cols) - s["logit_vals"]).abs().max()))
        nz = s["probs"] > 0
        distinct = torch.ones_like(nz)
        # a rank is comparable only if its probability differs from both neighbours by more than fp noise
        p = s["probs"]
        rel = (p[:, :-1] - p[:, 1:]) / p[:, :-1].clamp_min(1e-30)
        close = rel < 1e-3
        distinct[:, :-1] &= ~close
        distinct[:, 1:] &= ~close
        cmp = nz & distinct
        m["topk_id_mismatch"] += int((tr["idxs"][cmp] != s["idxs"][cmp]).sum())
        same = tr["idxs"] == s["idxs"]
        if bool((same & nz).any()):
            pr = ((tr["probs"] - s["probs"]).abs() / s["probs"].clamp_min(1e-30))[same & nz]
            m["prob_relerr"] = max(m["prob_relerr"], float(pr.max()))
        has_next = si + 1 < len(g["steps"]) and g["steps"][si + 1]["pos"] != pos
        for b in range(inp.shape[0]):
            ref_ids, got_ids = s["idxs"][b].tolist(), tr["idxs"][b].tolist()
            where = {}
            for k, v in enumerate(ref_ids):
                where.setdefault(v, k)  # masked candidates repeat id 0: first slot
            pairs = [(k, where[v]) for k, v in enumerate(got_ids) if v in where]
            if pairs:
                gk = torch.tensor([a for a, _ in pairs])
                rk = torch.tensor([c for _, c in pairs])
                if eng.precision != "certified":  # certified: only re-scored candidates carry exact cosines
                    m["clip_ref_err"] = max(m["clip_ref_err"], float((tr["clip_ref"][b][gk] - s["clip_ref"][b][rk]).abs().max()))
                    m["cos_checked"] += len(pairs)
            if sorted(ref_ids) != sorted(got_ids):
                continue
            if bool(same[b].all()) and eng.precision != "certified":
                m["clip_score_err"] = max(m["clip_score_err"], float((tr["clip_score"][b] - s["clip_score"][b]).abs().max()))
            # the reference's winner is visible in the next recorded step's inp
            if has_next:
                nxt = int(g["steps"][si + 1]["inp"][b, pos])
                got = int(inp[b, pos])
                m["winner_checked"] += 1
                if got != nxt:
                    m["winner_mismatch"] += 1
                    top2 = tr["final"][b].topk(min(2, K)).values
                    marg = float(top2[0] - top2[-1])
                    m["min_margin_at_mismatch"] = marg if m["min_margin_at_mismatch"] is None else min(
                        m["min_margin_at_mismatch"], marg)
                elif gamma is None:  # the reported cosine of the winner (gen_utils.py:80)
                    rw = int((0.02 * s["probs"][b] + 2.0 * s["clip_score"][b]).argmax())
                    m["winner_cos_err"] = max(m["winner_cos_err"], float((clip_ref[b] - s["clip_ref"][b][rw]).abs()))
    return m
