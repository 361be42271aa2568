"""CLIP ViT-B/32 image tower -> projected image embeddings, f32[B,512] (clip/clip.py:48-62;
HF:models/clip/modeling_clip.py:676-686).  Runs ONCE per generate_caption call, before the Gibbs loop; it is
the step before the hot path and is plain fp32 torch on the GPU for now --
not one of this repo's kernels and not counted in `gpu_launches`."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


@torch.no_grad()
def image_embeds(sd: Dict[str, torch.Tensor], pixel_values: torch.Tensor, heads: int = 12, eps: float = 1e-5):
    e = "vision_model.embeddings."
    B = pixel_values.shape[0]
    patch = sd[e + "patch_embedding.weight"]
    ps = patch.shape[-1]
    g = pixel_values.shape[-1] // ps
    # patch embedding as an fp32 matmul over unfolded patches (cuDNN convolutions default to TF32)
    cols = pixel_values.reshape(B, 3, g, ps, g, ps).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * ps * ps)
    x = cols @ patch.reshape(patch.shape[0], -1).t()
    x = torch.cat([sd[e + "class_embedding"].expand(B, 1, -1), x], dim=1) + sd[e + "position_embedding.weight"]
    H = x.shape[-1]
    ln = lambda t, n: F.layer_norm(t, (H,), sd[n + ".weight"], sd[n + ".bias"], eps)
    x = ln(x, "vision_model.pre_layrnorm")
    i = 0
    while f"vision_model.encoder.layers.{i}.layer_norm1.weight" in sd:
        p = f"vision_model.encoder.layers.{i}."
        lin = lambda t, n: F.linear(t, sd[p + n + ".weight"], sd[p + n + ".bias"])
        h = ln(x, p + "layer_norm1")
        N, T, _ = h.shape
        sh = lambda t: t.view(N, T, heads, H // heads).transpose(1, 2)
        q, k, v = sh(lin(h, "self_attn.q_proj")), sh(lin(h, "self_attn.k_proj")), sh(lin(h, "self_attn.v_proj"))
        a = torch.softmax((q @ k.transpose(-1, -2)) * (H // heads) ** -0.5, dim=-1) @ v
        x = x + lin(a.transpose(1, 2).reshape(N, T, H), "self_attn.out_proj")
        h = lin(ln(x, p + "layer_norm2"), "mlp.fc1")
        x = x + lin(h * torch.sigmoid(1.702 * h), "mlp.fc2")
        i += 1
    pooled = ln(x[:, 0], "vision_model.post_layernorm")
    return F.linear(pooled, sd["visual_projection.weight"])
