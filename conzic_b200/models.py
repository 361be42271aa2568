"""Weight holders with the surface the reference's entry points touch (`run.py:134-141`): something that
can be `.eval()`-ed, moved `.to(device)`, and handed to `generate_caption` as `model`.  They hold Hugging Face
state dicts only; all arithmetic happens in libconzic.so through `conzic_b200.runtime`."""
from __future__ import annotations

from typing import Dict

import torch


class BertMLM:
    """Stands where `AutoModelForMaskedLM.from_pretrained(...)` stands in the reference."""

    def __init__(self, state_dict: Dict[str, torch.Tensor]):
        self._sd = state_dict
        self.device = torch.device("cpu")

    @classmethod
    def from_pretrained(cls, name: str):
        from transformers import AutoModelForMaskedLM  # needs the checkpoint in the local HF cache
        return cls(AutoModelForMaskedLM.from_pretrained(name).state_dict())

    def state_dict(self):
        return self._sd

    def eval(self):
        return self

    def to(self, device):
        self.device = torch.device(device)
        return self

    def __call__(self, inp):
        """`model(inp).logits` for every position (gen_utils.py:69) -- L row evaluations; the generation loops
        never call this, they ask the engine for the one row they need."""
        from . import runtime
        eng = runtime.engine_for(self, None)
        rows = [eng.bert_mlm_row(inp, p) for p in range(inp.shape[1])]
        return type("MaskedLMOutput", (), {"logits": torch.stack(rows, dim=1)})()
