"""Conventional ids and constants of bert-base-uncased / CLIP ViT-B/32, used when a duck-typed tokenizer or
checkpoint does not state them (HF:models/bert/configuration_bert.py, HF:models/clip/configuration_clip.py)."""
PAD_ID, UNK_ID, CLS_ID, SEP_ID, MASK_ID = 0, 100, 101, 102, 103
SPECIAL_IDS = (PAD_ID, UNK_ID, CLS_ID, SEP_ID, MASK_ID)
DOT_ID = 1012
CLIP_BOS, CLIP_EOS = 49406, 49407
BERT_LN_EPS = 1e-12
CLIP_LN_EPS = 1e-5
