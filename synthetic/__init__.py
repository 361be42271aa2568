"""Synthetic weights, vocabularies, tokenizers and images for tests, tools and bench.py (no checkpoints or
vocabulary files exist offline).  Not imported by the product package except by `cli.py --synthetic`."""
