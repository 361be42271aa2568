"""`CLIP` with the reference wrapper's public methods (clip/clip.py:6-102 of the reference), backed by
libconzic.so for everything on the text side."""
from __future__ import annotations

from typing import Dict, Optional

import torch



class CLIP:
    def __init__(self, model_name: Optional[str] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 tokenizer=None, processor=None):
        """`CLIP(model_name)` loads a Hugging Face checkpoint like the reference (needs it in the local cache);
        `CLIP(state_dict=..., tokenizer=..., processor=...)` takes the pieces directly."""
        if state_dict is None:
            from transformers import CLIPModel, CLIPProcessor, CLIPTokenizer
            print("Initializing CLIP model...")
            state_dict = CLIPModel.from_pretrained(model_name).state_dict()
            processor = CLIPProcessor.from_pretrained(model_name)
            tokenizer = CLIPTokenizer.from_pretrained(model_name)
            print("CLIP model initialized.")
        self._sd = state_dict
        self.tokenizer = tokenizer
        self.processor = processor
        self.device = torch.device("cpu")

    # nn.Module-ish surface used by run.py / demo.py
    def state_dict(self):
        return self._sd

    def eval(self):
        return self

    def to(self, device):
        self.device = torch.device(device)
        return self

    def _engine(self):
        from .. import runtime
        return runtime.engine_for(None, self)

    # ---- image side: once per call, before the hot loop (clip/clip.py:48-62)
    @torch.no_grad()
    def compute_image_representation_from_image_instance(self, image):
        """PIL image(s) / uint8 arrays -> image embeddings.  With the Hugging Face CLIPImageProcessor in its standard
        configuration the pre-processing (antialiased bicubic resize, centre crop, normalise) runs on the device too
        (conzic_image_preprocess); any other processor is called as the reference calls it (clip/clip.py:55-58)."""
        from .. import imageproc
        cfg = imageproc.processor_config(self.processor)
        if cfg is not None:
            items = list(image) if isinstance(image, (list, tuple)) else [image]
            arrays = [imageproc.to_uint8_hwc(im) for im in items]
            if all(a is not None for a in arrays):
                eng = self._engine()
                return eng.image_encode(eng.preprocess_images(arrays, cfg))
        pixel_values = self.processor(images=image, return_tensors="pt")["pixel_values"]
        return self.compute_image_representation_from_pixels(pixel_values)

    @torch.no_grad()
    def compute_image_representation_from_pixels(self, pixel_values):
        """Vision tower on the engine's own sm_100a kernels (conzic_clip_image_encode)."""
        return self._engine().image_encode(pixel_values)

    @torch.no_grad()
    def compute_image_representation_from_image_path(self, image_path):
        from PIL import Image
        return self.compute_image_representation_from_image_instance(Image.open(image_path))

    # ---- text side (clip/clip.py:64-102)
    def tokenize_texts(self, text_list):
        """CLIP ids int64[N, T] of the texts, BOS ... EOS, right padded with EOS, truncated to 77 (clip/clip.py:71-72)."""
        t = self.tokenizer(text_list, padding=True, return_tensors="pt",
                           max_length=self.tokenizer.max_len_single_sentence + 2, truncation=True)
        return t["input_ids"]

    def compute_text_representation(self, text_list):
        eng = self._engine()
        return eng.clip_text_encode(self.tokenize_texts(text_list).to(eng.device))

    def compute_image_text_similarity_via_embeddings(self, image_embeds, text_embeds):
        return self._engine().image_text_similarity(image_embeds, text_embeds)

    def compute_image_text_similarity_via_raw_text(self, image_embeds, text_list):
        return self.compute_image_text_similarity_via_embeddings(image_embeds, self.compute_text_representation(text_list))
