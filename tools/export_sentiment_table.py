#!/usr/bin/env python
"""Build the per-vocabulary sentiment table that `control_gen_utils` takes instead of calling SentiWordNet on
every candidate caption (sentiments_classifer.py:9-33 of the reference).  Run this ONCE on a machine that has
NLTK with the `averaged_perceptron_tagger`, `wordnet` and `sentiwordnet` corpora and the BERT vocabulary; the
resulting tensor travels with the checkpoint (`run.py --sentiment_table table.pt`).

    python tools/export_sentiment_table.py --lm_model bert-base-uncased --out sentiment_table.pt

table[v] = the reference's score of the one-word text made of vocabulary entry v:
    tag the word, map the Penn tag to a WordNet class (n / v / a / r, anything else ''), and average
    pos_score - neg_score over its SentiWordNet synsets (0 when there are none).
The reference tags each word inside its caption, so a word whose tag depends on context can score differently
there; the table is exact for the words SentiWordNet scores the same under every tag the tagger gives them, and an
approximation otherwise.  '##' word pieces and special tokens score 0.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# sentiments_classifer.py:18-22
TAG_MAP = {"NN": "n", "NNP": "n", "NNPS": "n", "NNS": "n", "UH": "n",
           "VB": "v", "VBD": "v", "VBG": "v", "VBN": "v", "VBP": "v", "VBZ": "v",
           "JJ": "a", "JJR": "a", "JJS": "a",
           "RB": "r", "RBR": "r", "RBS": "r", "RP": "r", "WRB": "r"}


def word_score(word, pos_tag, senti_synsets) -> float:
    """Score of a one-word text, following sentiments_classifer.py:14-30."""
    tagged = pos_tag([word])
    total = 0.0
    for w, t in tagged:
        syn = list(senti_synsets(w, TAG_MAP.get(t, "")))
        if syn:
            total += sum(x.pos_score() - x.neg_score() for x in syn) / len(syn)
    return total


def build_table(tokens, pos_tag, senti_synsets, special=("[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]")) -> torch.Tensor:
    table = torch.zeros(len(tokens), dtype=torch.float32)
    for v, tok in enumerate(tokens):
        if tok in special or tok.startswith("##") or tok.startswith("[unused"):
            continue
        table[v] = word_score(tok, pos_tag, senti_synsets)
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lm_model", default="bert-base-uncased")
    ap.add_argument("--out", default="sentiment_table.pt")
    a = ap.parse_args()
    try:
        from nltk import pos_tag
        from nltk.corpus import sentiwordnet
    except ImportError as e:
        raise SystemExit("this tool needs nltk with the averaged_perceptron_tagger, wordnet and sentiwordnet corpora "
                         f"({e}); it is meant to run once, off line, next to the checkpoint") from e
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(a.lm_model)
    tokens = tok.convert_ids_to_tokens(list(range(tok.vocab_size)))
    table = build_table(tokens, pos_tag, sentiwordnet.senti_synsets)
    torch.save(table, a.out)
    nz = int((table != 0).sum())
    print(f"wrote {a.out}: {len(tokens)} entries, {nz} non-zero, range [{float(table.min()):.3f}, {float(table.max()):.3f}]")


if __name__ == "__main__":
    main()
