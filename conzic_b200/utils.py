"""Helpers with the reference's names and meaning (utils.py:8-59 of the reference): seeds, the initial
"prompt + [MASK]*n" ids, the '.'-only-at-the-end rule and a logger.  Host-side Python, no device work."""
from __future__ import annotations

import logging
import os
import random
import sys
import time

import numpy as np
import torch


def create_logger(folder, filename):
    """A 'ConZIC' logger with a stream handler and a file handler under `folder` (utils.py:8-35).
    colorlog is used when it is installed; plain formatting otherwise."""
    os.makedirs(folder, exist_ok=True)
    logger = logging.getLogger("ConZIC")
    logger.setLevel(logging.INFO)
    logger.handlers = []
    fmt = "%(asctime)s %(levelname)s: %(message)s"
    try:
        import colorlog  # type: ignore
        stream_fmt = colorlog.ColoredFormatter("%(log_color)s" + fmt)
    except Exception:  # noqa: BLE001
        stream_fmt = logging.Formatter(fmt)
    sh = logging.StreamHandler(sys.stdout)
    sh.setFormatter(stream_fmt)
    fh = logging.FileHandler(os.path.join(folder, filename))
    fh.setFormatter(logging.Formatter(fmt))
    logger.addHandler(sh)
    logger.addHandler(fh)
    return logger


def set_seed(seed):
    """Seeds python, numpy and torch exactly like the reference so visiting orders reproduce (utils.py:37-44)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def get_init_text(tokenizer, seed_text, max_len, batch_size=1):
    """ids of `seed_text` followed by max_len mask tokens, replicated batch_size times (utils.py:46-51)."""
    ids = tokenizer.encode(seed_text + tokenizer.mask_token * max_len)
    return [ids for _ in range(batch_size)]


def update_token_mask(tokenizer, token_mask, max_len, index):
    """'.' may only be generated at the last caption position; mutates token_mask in place (utils.py:53-59)."""
    token_mask[:, tokenizer.vocab["."]] = 1 if index == max_len - 1 else 0
    return token_mask


def format_output(sample_num, FinalCaption, BestCaption):
    """(final, best): the first `sample_num` captions of each list joined by newlines, at most five -- the two
    strings the reference's Gradio front end shows (utils.py:61-74)."""
    n = min(max(int(sample_num), 1), 5)
    return "\n".join(f"{c}" for c in FinalCaption[:n]), "\n".join(f"{c}" for c in BestCaption[:n])
