#!/usr/bin/env python
"""Measures the error bound the certified argmax rests on (conzic_config.cert_dcos): the distribution of
|cos(bf16 tower) - cos(bf16x3 tower)| over every candidate of free-running Gibbs steps at BASELINE config 2 sizes,
on the production step path (shared-prefix layout, LayerNorm in the GEMM epilogues), and what a given bound costs
(how many candidates per image survive round 1 of cert_ops.cu).

The chain is driven by the bf16x3 engine; the bf16 engine is teacher-forced on the same ids and the same BERT
logits (conzic_gibbs_step's logits_in), so both towers score exactly the same candidate captions.

    python tools/cert_bound.py [--images 64] [--sweeps 5] [--K 200] [--len 10] > gpurun_out/cert_bound.json
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthetic import synth  # noqa: E402


def survivor_mask(tr16, hi, lo, alpha=0.02, beta=2.0):
    """Round 1 of cert_ops.cu in torch with the asymmetric bounds (logit units): True = cannot be ruled out."""
    a = tr16["clip_ref"].double() * 100.0
    p = tr16["probs"].double()
    m = a.max(dim=1, keepdim=True).values
    e = torch.exp(a - m)
    Z = e.sum(dim=1, keepdim=True)
    f = alpha * p + beta * e / Z
    w = f.argmax(dim=1, keepdim=True)
    ew, Ew = e.gather(1, w), (alpha * p).gather(1, w)
    d = ew * math.exp(-hi) - e * math.exp(lo)
    x = torch.where(d >= 0, d / (Z * math.exp(lo)), d / (Z * math.exp(-hi)))
    alive = ~((Ew - alpha * p) + beta * x > 1e-5)
    alive.scatter_(1, w, True)
    return alive


def survivors(tr3, tr16, eps_logit, alpha=0.02, beta=2.0):
    """Round 1 of the certified argmax in torch, on the bf16 cosines: candidates whose lower bound does not clear 0."""
    a = tr16["clip_ref"].double() * 100.0
    p = tr16["probs"].double()
    m = a.max(dim=1, keepdim=True).values
    e = torch.exp(a - m)
    Z = e.sum(dim=1, keepdim=True)
    f = alpha * p + beta * e / Z
    w = f.argmax(dim=1, keepdim=True)
    ew, Ew = e.gather(1, w), (alpha * p).gather(1, w)
    d = ew * math.exp(-eps_logit) - e * math.exp(eps_logit)
    x = torch.where(d >= 0, d / (Z * math.exp(eps_logit)), d / (Z * math.exp(-eps_logit)))
    lb = (Ew - alpha * p) + beta * x
    alive = ~(lb > 1e-5)
    alive.scatter_(1, w, True)
    return alive.sum(dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--sweeps", type=int, default=5)
    ap.add_argument("--K", type=int, default=200)
    ap.add_argument("--len", type=int, default=10)
    ap.add_argument("--order", default="sequential")
    args = ap.parse_args()
    from conzic_b200.engine import Engine
    bert_sd, clip_sd = synth.make_bert_state_dict(0), synth.make_clip_state_dict(0, vision=True)
    e3 = Engine(bert_sd, clip_sd, precision="bf16x3")
    e16 = Engine(bert_sd, clip_sd, precision="bf16")
    off, tok = synth.build_bert2clip_table(False)
    e3.set_bert2clip(off, tok)
    e16.set_bert2clip(off, tok)
    B, n, K = args.images, args.len, args.K
    # the bench's inputs: uint8 images through the device CLIPImageProcessor
    import numpy as np
    from transformers import CLIPImageProcessor
    from conzic_b200 import imageproc
    cfg = imageproc.processor_config(CLIPImageProcessor())
    raw = torch.from_numpy(np.stack([synth.make_uint8_image(i) for i in range(B)])).cuda()
    pix = e3.preprocess_uint8(raw, cfg)
    img = e3.image_encode(pix)
    img16 = e16.image_encode(pix)
    icos = torch.nn.functional.cosine_similarity(img, img16, dim=-1)
    tokz = synth.SynthBertTokenizer()
    inp = torch.tensor([tokz.encode(synth.SYNTH_PROMPT + "[MASK]" * n)] * B).cuda()
    holds = [False] + [True] * 3 + [False] * (n + 1)
    order = list(range(n))
    if args.order == "shuffle":
        import random
        random.seed(42)
        random.shuffle(order)
    eps_grid = [1e-3, 1.5e-3, 2e-3, 3e-3, 4e-3, 6e-3, 8e-3]
    all_d, all_sd, all_cen, flips, steps = [], [], [], 0, 0
    r_all, r_rest = [], []  # sum_j exp(exact logit) / sum_j exp(bf16 logit) over all candidates / over the ones ruled out
    surv = {e: [] for e in eps_grid}
    per_step = []
    for it in range(args.sweeps):
        for ii in order:
            pos = 4 + ii
            tm3, tm16 = synth.make_token_mask("cuda"), synth.make_token_mask("cuda")
            masked = inp.clone()
            masked[:, pos] = synth.MASK_ID
            logits = e3.bert_mlm_row_padded(masked, pos)
            i3, i16 = inp.clone(), inp.clone()
            kw = dict(logits_in=logits, trace=True)
            before, after = sum(holds[:pos]), sum(holds[pos + 1:])
            _, _, t3 = e3.gibbs_step(i3, tm3, img, pos, ii == n - 1, K, 0.1, 0.02, 2.0, before, after, **kw)
            _, _, t16 = e16.gibbs_step(i16, tm16, img, pos, ii == n - 1, K, 0.1, 0.02, 2.0, before, after, **kw)
            torch.cuda.synchronize()
            assert torch.equal(t3["idxs"], t16["idxs"])
            sd = t16["clip_ref"] - t3["clip_ref"]  # signed: the softmax only sees differences between candidates
            d = sd.abs()
            all_d.append(d.flatten().cpu())
            all_sd.append(sd.flatten().cpu())
            cen = sd - sd.mean(dim=1, keepdim=True)
            all_cen.append(cen.flatten().cpu())
            a16, a3 = t16["clip_ref"].double() * 100.0, t3["clip_ref"].double() * 100.0
            mref = a16.max(dim=1, keepdim=True).values
            x16, x3 = torch.exp(a16 - mref), torch.exp(a3 - mref)
            r_all.append((x3.sum(1) / x16.sum(1)).cpu())
            rest = ~survivor_mask(t16, 0.16, 0.071)
            has = rest.any(dim=1)
            r_rest.append(((x3 * rest).sum(1)[has] / (x16 * rest).sum(1)[has]).cpu())
            nf = int((i3[:, pos] != i16[:, pos]).sum())
            flips += nf
            steps += 1
            for e in eps_grid:
                surv[e].append(survivors(t3, t16, 100.0 * e).cpu())
            per_step.append(dict(sweep=it, ii=ii, max_dcos=float(d.max()), winner_flips=nf))
            holds[pos] = True
            inp = i3
    d = torch.cat(all_d)
    qs = [0.5, 0.9, 0.99, 0.999, 0.9999, 0.99999]
    ds = d.sort().values
    quant = {str(q): float(ds[min(int(q * ds.numel()), ds.numel() - 1)]) for q in qs}
    sd, cen = torch.cat(all_sd), torch.cat(all_cen)
    signed = dict(mean=float(sd.mean()), std=float(sd.std()), min=float(sd.min()), max=float(sd.max()),
                  max_abs_minus_global_mean=float((sd - sd.mean()).abs().max()),
                  max_abs_minus_image_mean=float(cen.abs().max()), std_minus_image_mean=float(cen.std()))
    ra, rr = torch.cat(r_all), torch.cat(r_rest)
    zratio = dict(all=dict(min=float(ra.min()), max=float(ra.max()), mean=float(ra.mean())),
                  ruled_out=dict(min=float(rr.min()), max=float(rr.max()), mean=float(rr.mean()), n=int(rr.numel())))
    rep = dict(signed_error=signed, softmax_denominator_ratio_exact_over_bf16=zratio, workload=dict(images=B, sweeps=args.sweeps, K=K, sentence_len=n, order=args.order, steps=steps,
                             candidates=int(d.numel())),
               max_dcos=float(d.max()), mean_dcos=float(d.mean()), quantiles=quant,
               image_embed_min_cos_bf16_vs_x3=float(icos.min()),
               bf16_winner_flips=flips, decisions=steps * B,
               survivors_per_image={str(e): dict(mean=float(torch.cat(surv[e]).float().mean()),
                                                 max=int(torch.cat(surv[e]).max()),
                                                 frac_images_multi=float((torch.cat(surv[e]) > 1).float().mean()))
                                    for e in eps_grid},
               per_step=per_step)
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
