import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def synth_weights():
    """Synthetic HF-keyed state dicts, generated once per session (about 10 s)."""
    from synthetic import synth
    cache = {}

    def get(kind, peaked=False):
        key = (kind, peaked)
        if key not in cache:
            cache[key] = (synth.make_bert_state_dict(0, peaked=peaked) if kind == "bert"
                          else synth.make_clip_state_dict(0))
        return cache[key]
    return get


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
