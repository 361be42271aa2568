"""Command lines of the reference's two entry points, `demo.py` (one image, demo.py:105-152) and `run.py`
(an image directory in batches with JSON results, run.py:114-222), on top of the libconzic engine.

Flag names and defaults are the reference's (run.py:15-76 / demo.py:15-76).  Additions, all optional:
  --synthetic            random-init bert-base / CLIP ViT-B/32 shaped weights, synthetic tokenizers and images
                         (no checkpoint, vocabulary or image file is needed: this image has no network)
  --synthetic_images N   number of synthetic images for run.py --synthetic
  --precision            bf16 (throughput) or bf16x3 (3-pass split operands, reference-identical token ids)
  --sentiment_table F    torch file with f32[V] per-vocabulary control scores (replaces the NLTK scorer)
Under torchrun (`WORLD_SIZE > 1`) run.py shards the batches over the ranks (one process per GPU) and rank 0
writes the result files; visiting orders are drawn on every rank so results do not depend on the rank count.
"""
from __future__ import annotations

import argparse
import json
import os
import time

import torch

from . import control_gen_utils, dist, gen_utils


def _synth():
    """The offline fixtures behind --synthetic (repo-level package `synthetic`, not part of the product package)."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("synthetic.synth")
from .clip.clip import CLIP
from .models import BertMLM
from .utils import create_logger, set_seed

_POS_TEMPLATE = [["DET"], ["ADJ", "NOUN"], ["NOUN"], ["VERB"], ["VERB"], ["ADV"], ["ADP"], ["DET", "NOUN"],
                 ["NOUN"], ["NOUN", "."], [".", "NOUN"], [".", "NOUN"]]


def get_args(entry: str, argv=None):
    """entry = "demo" or "run": the two scripts differ in three defaults only."""
    demo = entry == "demo"
    p = argparse.ArgumentParser(prog=f"{entry}.py")
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--batch_size", type=int, default=1 if demo else 2)
    p.add_argument("--device", type=str, default="cuda", choices=["cuda", "cpu"])
    p.add_argument("--run_type", default="controllable", nargs="?", choices=["caption", "controllable"])
    p.add_argument("--prompt", default="Image of a", type=str)
    p.add_argument("--order", default="shuffle", nargs="?", choices=["sequential", "shuffle", "span", "random"],
                   help="Generation order of text")
    p.add_argument("--control_type", default="sentiment", nargs="?", choices=["sentiment", "pos"])
    p.add_argument("--pos_type", type=list, default=_POS_TEMPLATE, help="predefined part-of-speech template")
    p.add_argument("--sentiment_type", default="positive", nargs="?", choices=["positive", "negative"])
    p.add_argument("--samples_num", default=2, type=int)
    p.add_argument("--sentence_len", type=int, default=10)
    p.add_argument("--candidate_k", type=int, default=200)
    p.add_argument("--alpha", type=float, default=0.02, help="weight for fluency")
    p.add_argument("--beta", type=float, default=2.0, help="weight for image-matching degree")
    p.add_argument("--gamma", type=float, default=5.0, help="weight for controllable degree")
    p.add_argument("--lm_temperature", type=float, default=0.1)
    p.add_argument("--num_iterations", type=int, default=10, help="predefined iterations for Gibbs Sampling")
    p.add_argument("--lm_model", type=str, default="bert-base-uncased")
    p.add_argument("--match_model", type=str,
                   default="openai/clip-vit-base-patch32" if demo else "clip-vit-base-patch32")
    p.add_argument("--caption_img_path", type=str, default="./examples/girl.jpg" if demo else "./examples/")
    p.add_argument("--stop_words_path", type=str, default="stop_words.txt")
    p.add_argument("--add_extra_stopwords", type=list, default=[])
    # additions
    p.add_argument("--synthetic", action="store_true")
    p.add_argument("--synthetic_images", type=int, default=8)
    p.add_argument("--precision", default=None, choices=["bf16", "bf16x3"])
    p.add_argument("--sentiment_table", type=str, default=None)
    p.add_argument("--results_dir", type=str, default="results")
    return p.parse_args(argv)


def _run_label(args):
    label = "caption" if args.run_type == "caption" else args.control_type
    return args.sentiment_type if label == "sentiment" else label


def _make_logger(args, prefix):
    label = _run_label(args)
    stamp = time.strftime("%Y-%m-%d-%H-%M-%S", time.localtime())
    name = (f"{prefix}{label}_{args.order}_len{args.sentence_len}_topk{args.candidate_k}_alpha{args.alpha}"
            f"_beta{args.beta}_gamma{args.gamma}_lmtemp{args.lm_temperature}_{stamp}.log")
    logger = create_logger("logger", name)
    logger.info(f"Generating order:{args.order}")
    logger.info(f"Run type:{label}")
    logger.info(args)
    return logger


def load_models(args):
    """(lm_model, lm_tokenizer, clip, sentiment_table or None) on the engine's device."""
    if args.device != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("conzic_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
    if args.precision:
        os.environ["CONZIC_PRECISION"] = args.precision
    _, local_rank, _ = dist.env_world()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    table = None
    if args.synthetic:
        lm_model = BertMLM(_synth().make_bert_state_dict(0))
        lm_tokenizer = _synth().SynthBertTokenizer()
        clip = CLIP(state_dict=_synth().make_clip_state_dict(0), tokenizer=_synth().SynthCLIPTokenizer(),
                    processor=_synth().SynthProcessor())
        table = _synth().make_sentiment_table()
        control_gen_utils.set_pos_tagger(_synth().synth_pos_tagger)  # --control_type pos without NLTK
    else:
        from transformers import AutoTokenizer
        lm_model = BertMLM.from_pretrained(args.lm_model)
        lm_tokenizer = AutoTokenizer.from_pretrained(args.lm_model)
        clip = CLIP(args.match_model)
    if args.sentiment_table:
        table = torch.load(args.sentiment_table)
    lm_model.eval()
    clip.eval()
    return lm_model.to(dev), lm_tokenizer, clip.to(dev), table, dev


def build_token_mask(args, tokenizer, dev):
    """ones(1, V) with the stop words zeroed (run.py:143-152).  --synthetic uses the id-range rule of
    SURVEY.md 8(d) (ids < 1996 are specials / [unused] / punctuation / digits in bert-base-uncased)."""
    if args.synthetic and not os.path.exists(args.stop_words_path):
        return _synth().make_token_mask(dev)
    if not os.path.exists(args.stop_words_path):
        raise FileNotFoundError(f"--stop_words_path {args.stop_words_path!r} not found: point it at the reference checkout's "
                                "stop_words.txt (the list is the reference's data file and is not duplicated here)")
    with open(args.stop_words_path, "r", encoding="utf-8") as fh:
        words = [w.rstrip("\n") for w in fh.readlines()] + list(args.add_extra_stopwords)
    mask = torch.ones((1, tokenizer.vocab_size))
    for i in tokenizer.convert_tokens_to_ids(words):
        mask[0, i] = 0
    return mask.to(dev)


def _generate(args, names, images, lm_model, lm_tokenizer, clip, token_mask, logger, table):
    common = dict(prompt=args.prompt, batch_size=len(names), max_len=args.sentence_len, top_k=args.candidate_k,
                  temperature=args.lm_temperature, max_iter=args.num_iterations, alpha=args.alpha, beta=args.beta,
                  generate_order=args.order)
    if args.run_type == "caption":
        return gen_utils.generate_caption(names, lm_model, clip, lm_tokenizer, images, token_mask, logger, **common)
    if args.run_type == "controllable":
        return control_gen_utils.control_generate_caption(
            names, lm_model, clip, lm_tokenizer, images, token_mask, logger, gamma=args.gamma,
            ctl_type=args.control_type, style_type=args.sentiment_type, pos_type=args.pos_type,
            sentiment_table=table, **common)
    raise Exception("run_type must be caption or controllable!")


def demo_main(argv=None):
    args = get_args("demo", argv)
    set_seed(args.seed)
    logger = _make_logger(args, "demo_")
    lm_model, lm_tokenizer, clip, table, dev = load_models(args)
    token_mask = build_token_mask(args, lm_tokenizer, dev)
    logger.info(f"Processing: {args.caption_img_path}")
    if args.synthetic:
        image, name = _synth().make_pixel_values(0).unsqueeze(0), ["synthetic0.jpg"]
    else:
        from PIL import Image
        image, name = Image.open(args.caption_img_path).convert("RGB"), [args.caption_img_path.split("/")[-1]]
    out = None
    with torch.no_grad():
        for sample_id in range(args.samples_num):
            logger.info(f"Sample {sample_id}: ")
            out = _generate(args, name, image, lm_model, lm_tokenizer, clip, token_mask, logger, table)
    return out


def _results_dir(args, sample_id):
    head = f"caption_{args.order}" if args.run_type == "caption" else f"{_run_label(args)}_{args.order}"
    return os.path.join(args.results_dir, "%s_len%d_topk%d_alpha%.3f_beta%.3f_gamma%.3f_lmTemp%.3f" % (
        head, args.sentence_len, args.candidate_k, args.alpha, args.beta, args.gamma, args.lm_temperature),
        "sample_%d" % sample_id)


def run_main(argv=None):
    args = get_args("run", argv)
    set_seed(args.seed)
    rank, _, world = dist.init() if int(os.environ.get("WORLD_SIZE", 1)) > 1 else (0, 0, 1)
    logger = _make_logger(args, "" if world == 1 else f"rank{rank}_")
    lm_model, lm_tokenizer, clip, table, dev = load_models(args)
    token_mask = build_token_mask(args, lm_tokenizer, dev)
    if args.synthetic:
        names_all = [f"synthetic{i}.jpg" for i in range(args.synthetic_images)]
        load = lambda idx: torch.stack([_synth().make_pixel_values(i) for i in idx])
    else:
        from PIL import Image
        names_all = os.listdir(args.caption_img_path)  # the reference's order (run.py:159)
        load = lambda idx: [Image.open(os.path.join(args.caption_img_path, names_all[i])).convert("RGB") for i in idx]
    batches = dist.batches_of(len(names_all), args.batch_size)  # drop_last=True like the reference's DataLoader
    written = []

    def run_batch(sample_id, bi, idx):
        logger.info(f"The {bi+1}-th batch:")
        names = [names_all[i] for i in idx]
        with torch.no_grad():
            texts, _ = _generate(args, names, load(idx), lm_model, lm_tokenizer, clip, token_mask, logger, table)
        return names, texts

    def skip_batch(sample_id, bi, idx):  # another rank's batch: consume the RNG exactly like a call would
        dist.consume_order_rng(args.order if args.run_type == "caption" or args.order == "sequential" else "shuffle",
                               args.sentence_len, args.num_iterations)

    for sample_id in range(args.samples_num):
        logger.info(f"Sample {sample_id+1}: ")
        mine = dist.run_sharded(1, batches, lambda s, bi, idx: run_batch(sample_id, bi, idx),
                                lambda s, bi, idx: skip_batch(sample_id, bi, idx), rank, world)
        gathered = dist.gather_objects(mine, world)
        if rank != 0:
            continue
        n_lists = args.num_iterations + 1
        all_results = [dict() for _ in range(n_lists)]
        for part in gathered:
            for (_, bi), (names, texts) in sorted(part.items()):
                for it, cap in enumerate(texts[:n_lists]):
                    for jj, nm in enumerate(names):
                        all_results[it][nm.split(".")[0]] = cap[jj]
        save_dir = _results_dir(args, sample_id)
        os.makedirs(save_dir, exist_ok=True)
        for it, res in enumerate(all_results):
            fname = "best_clipscore.json" if it == n_lists - 1 else f"iter_{it}.json"
            with open(os.path.join(save_dir, fname), "w") as fh:
                json.dump(res, fh)
        written.append(save_dir)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return written
