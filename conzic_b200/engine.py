"""Python host side of libconzic.so: owns the device tensors, passes raw pointers through the C ABI.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed only; every compute call
below lands in a hand-written sm_100a kernel.  Nothing here falls back to torch ops or to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from . import defaults as _dims

PRECISIONS = {"bf16": _lib.PREC_BF16, "bf16x3": _lib.PREC_BF16X3, "certified": _lib.PREC_CERTIFIED}

SD = Dict[str, torch.Tensor]


def _bert_table(sd: SD, layers: int):
    e, p = "bert.embeddings.", "cls.predictions."
    names = [e + "word_embeddings.weight", e + "position_embeddings.weight", e + "token_type_embeddings.weight",
             e + "LayerNorm.weight", e + "LayerNorm.bias", p + "bias", p + "transform.dense.weight",
             p + "transform.dense.bias", p + "transform.LayerNorm.weight", p + "transform.LayerNorm.bias"]
    for i in range(layers):
        q = f"bert.encoder.layer.{i}."
        names += [q + "attention.self.query.weight", q + "attention.self.query.bias",
                  q + "attention.self.key.weight", q + "attention.self.key.bias",
                  q + "attention.self.value.weight", q + "attention.self.value.bias",
                  q + "attention.output.dense.weight", q + "attention.output.dense.bias",
                  q + "attention.output.LayerNorm.weight", q + "attention.output.LayerNorm.bias",
                  q + "intermediate.dense.weight", q + "intermediate.dense.bias",
                  q + "output.dense.weight", q + "output.dense.bias",
                  q + "output.LayerNorm.weight", q + "output.LayerNorm.bias"]
    return names


def _clip_table(sd: SD, layers: int):
    names = ["text_model.embeddings.token_embedding.weight", "text_model.embeddings.position_embedding.weight",
             "text_model.final_layer_norm.weight", "text_model.final_layer_norm.bias", "text_projection.weight"]
    for i in range(layers):
        q = f"text_model.encoder.layers.{i}."
        names += [q + "layer_norm1.weight", q + "layer_norm1.bias",
                  q + "self_attn.q_proj.weight", q + "self_attn.q_proj.bias",
                  q + "self_attn.k_proj.weight", q + "self_attn.k_proj.bias",
                  q + "self_attn.v_proj.weight", q + "self_attn.v_proj.bias",
                  q + "self_attn.out_proj.weight", q + "self_attn.out_proj.bias",
                  q + "layer_norm2.weight", q + "layer_norm2.bias",
                  q + "mlp.fc1.weight", q + "mlp.fc1.bias", q + "mlp.fc2.weight", q + "mlp.fc2.bias"]
    return names


def _count_layers(sd: SD, fmt: str) -> int:
    n = 0
    while fmt.format(n) in sd:
        n += 1
    return n


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    """One context per (BERT weights, CLIP text weights, device).  Mirrors what `model(inp).logits`,
    `generate_caption_step` and `CLIP.compute_text_representation / ..._similarity_via_embeddings` compute in the
    reference (gen_utils.py:64-81, clip/clip.py:64-98)."""

    def __init__(self, bert_sd: SD, clip_sd: SD, device="cuda:0", precision: str = "certified",
                 gemm_impl: str = "tcgen05", special_ids: Sequence[int] = _dims.SPECIAL_IDS,
                 dot_id: int = _dims.DOT_ID, clip_bos: int = _dims.CLIP_BOS, clip_eos: int = _dims.CLIP_EOS,
                 clip_chunk_rows: int = 0, cert_dcos: Optional[float] = None, cert_dcos_lo: Optional[float] = None,
                 cert_zratio: Optional[Sequence[float]] = None, cert_fcap: Optional[int] = None,
                 ln_standalone: Optional[bool] = None, pdl: Optional[bool] = None, wide_lsu: Optional[bool] = None,
                 lsu_out: Optional[bool] = None):
        """precision: "certified" (default; the reference's token ids at close to bf16 speed), "bf16x3" (everything
        in the fp32-grade split mode) or "bf16" (tolerance-only parity); see include/conzic.h.  The remaining switches
        are read once, here: cert_dcos / cert_dcos_lo / cert_fcap (CONZIC_CERT_DCOS / _LO / CONZIC_CERT_FCAP), ln_standalone
        (CONZIC_LN_STANDALONE=1), pdl (CONZIC_PDL=0), wide_lsu (CONZIC_WIDE_LSU=1: per-lane instead of TMA reduce-add
        epilogue of the N = 512 GEMM) and lsu_out (CONZIC_LSU_OUT=1: per-lane instead of TMA stores of the bf16 GEMM
        outputs) exist for A/B measurements."""
        if not torch.cuda.is_available():
            raise RuntimeError("conzic_b200.Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.index is None:  # clip.to("cuda") in the reference's scripts
            self.device = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        self.precision = precision
        nb = _count_layers(bert_sd, "bert.encoder.layer.{}.attention.self.query.weight")
        nc = _count_layers(clip_sd, "text_model.encoder.layers.{}.layer_norm1.weight")
        word = bert_sd["bert.embeddings.word_embeddings.weight"]
        tok = clip_sd["text_model.embeddings.token_embedding.weight"]
        cfg = _lib.Config()
        cfg.bert_layers, cfg.bert_hidden, cfg.bert_vocab = nb, word.shape[1], word.shape[0]
        cfg.bert_heads = word.shape[1] // 64
        cfg.bert_ffn = bert_sd["bert.encoder.layer.0.intermediate.dense.weight"].shape[0]
        cfg.bert_maxpos = bert_sd["bert.embeddings.position_embeddings.weight"].shape[0]
        cfg.bert_ln_eps = _dims.BERT_LN_EPS
        cfg.clip_layers, cfg.clip_hidden, cfg.clip_vocab = nc, tok.shape[1], tok.shape[0]
        cfg.clip_heads = tok.shape[1] // 64
        cfg.clip_ffn = clip_sd["text_model.encoder.layers.0.mlp.fc1.weight"].shape[0]
        cfg.clip_maxpos = clip_sd["text_model.embeddings.position_embedding.weight"].shape[0]
        cfg.clip_proj = clip_sd["text_projection.weight"].shape[0]
        cfg.clip_ln_eps = _dims.CLIP_LN_EPS
        cfg.pad_id, cfg.unk_id, cfg.cls_id, cfg.sep_id, cfg.mask_id = [int(x) for x in special_ids]
        cfg.dot_id, cfg.clip_bos, cfg.clip_eos = int(dot_id), int(clip_bos), int(clip_eos)
        cfg.precision = PRECISIONS[precision]
        cfg.gemm_impl = {"tcgen05": _lib.GEMM_TCGEN05, "simt_debug": _lib.GEMM_SIMT_DEBUG}[gemm_impl]
        cfg.clip_chunk_rows = int(clip_chunk_rows)
        env = os.environ.get
        cfg.cert_dcos = float(cert_dcos if cert_dcos is not None else env("CONZIC_CERT_DCOS", 0.0))
        cfg.cert_dcos_lo = float(cert_dcos_lo if cert_dcos_lo is not None else env("CONZIC_CERT_DCOS_LO", 0.0))
        if cert_zratio is None and env("CONZIC_CERT_ZRATIO"):
            cert_zratio = [float(v) for v in env("CONZIC_CERT_ZRATIO").split(",")]
        cfg.cert_zratio_lo, cfg.cert_zratio_hi = (float(cert_zratio[0]), float(cert_zratio[1])) if cert_zratio else (0.0, 0.0)
        cfg.cert_fcap = int(cert_fcap if cert_fcap is not None else env("CONZIC_CERT_FCAP", 0))
        if ln_standalone is None:
            ln_standalone = env("CONZIC_LN_STANDALONE", "0") == "1"
        if pdl is None:
            pdl = env("CONZIC_PDL", "1") != "0"
        if wide_lsu is None:
            wide_lsu = env("CONZIC_WIDE_LSU", "0") == "1"
        if lsu_out is None:
            lsu_out = env("CONZIC_LSU_OUT", "0") == "1"
        cfg.flags = ((_lib.FLAG_LN_STANDALONE if ln_standalone else 0) | (0 if pdl else _lib.FLAG_NO_PDL) |
                     (_lib.FLAG_WIDE_LSU if wide_lsu else 0) | (_lib.FLAG_LSU_OUT if lsu_out else 0))
        self.cfg = cfg
        self.V, self.D = cfg.bert_vocab, cfg.clip_proj
        self.ldl = (self.V + 3) & ~3
        self.mask_id = cfg.mask_id

        def table(names, sd):
            ts = [sd[n].detach().to(self.device, torch.float32).contiguous() for n in names]
            arr = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
            return ts, arr

        bt, barr = table(_bert_table(bert_sd, nb), bert_sd)
        ct, carr = table(_clip_table(clip_sd, nc), clip_sd)
        ctx = C.c_void_p()
        rc = self.lib.conzic_ctx_create(C.byref(cfg), barr, len(bt), carr, len(ct), self._stream(), C.byref(ctx))
        _lib.check(rc, "conzic_ctx_create")
        del bt, ct  # the context holds its own copies; create synchronised the stream
        self.ctx = ctx
        ls = clip_sd.get("logit_scale")
        self.logit_scale_exp = float(torch.as_tensor(ls).float().exp()) if ls is not None else 100.0
        self._ws = None
        self._vws = None
        self.vision_cfg = None
        self.has_table = False
        self.has_text_vocab = False
        if "vision_model.embeddings.patch_embedding.weight" in clip_sd:
            self.set_vision(clip_sd)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "ctx", None):
            torch.cuda.synchronize(self.device)
            self.lib.conzic_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def workspace(self, B: int, L: int, K: int) -> torch.Tensor:
        need = int(self.lib.conzic_workspace_bytes(self.ctx, B, L, K))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def launch_count(self) -> int:
        return int(self.lib.conzic_launch_count(self.ctx))

    CERT_STAT_NAMES = ("calls", "images", "rescored_candidates", "images_multi", "images_full_overflow", "images_full")

    def cert_stats(self) -> Dict[str, int]:
        """Certified-argmax bookkeeping since the engine was created (conzic_cert_stats); zeros in other modes."""
        arr = (C.c_uint64 * _lib.CERT_STATS)()
        _lib.check(self.lib.conzic_cert_stats(self.ctx, arr, _lib.CERT_STATS), "conzic_cert_stats")
        return {n: int(arr[i]) for i, n in enumerate(self.CERT_STAT_NAMES)}

    PROFILE_CATEGORIES = ("gemm", "attention", "layernorm", "embed", "topk", "assemble", "select", "misc",
                          "gemm_small")

    def profile(self, enable: bool):
        """CUDA-event timing of every launch by category (conzic_profile); off by default."""
        _lib.check(self.lib.conzic_profile(self.ctx, 1 if enable else 0), "conzic_profile")

    PROFILE_PHASES = ("other", "bert", "candidates", "tower", "select", "cert_rescore", "cert_full", "image")

    def profile_read_phases(self):
        """{phase: (milliseconds, launches)} of the same records, by step phase (conzic_profile_read 100 + p)."""
        out = {}
        for i, name in enumerate(self.PROFILE_PHASES):
            ms, work, n = C.c_double(), C.c_double(), C.c_int()
            _lib.check(self.lib.conzic_profile_read(self.ctx, 100 + i, C.byref(ms), C.byref(work), C.byref(n)),
                       "conzic_profile_read")
            out[name] = (ms.value, n.value)
        return out

    def profile_read(self):
        """{category: (milliseconds, work, launches)}; waits for the recorded events."""
        out = {}
        for i, name in enumerate(self.PROFILE_CATEGORIES):
            ms, work, n = C.c_double(), C.c_double(), C.c_int()
            _lib.check(self.lib.conzic_profile_read(self.ctx, i, C.byref(ms), C.byref(work), C.byref(n)),
                       "conzic_profile_read")
            out[name] = (ms.value, work.value, n.value)
        return out

    def set_bert2clip(self, off: torch.Tensor, tok: torch.Tensor):
        """CSR table BERT id -> CLIP BPE ids (int32).  See conzic_b200.tokens for how it is built."""
        off = off.to(self.device, torch.int32).contiguous()
        tok = tok.to(self.device, torch.int32).contiguous()
        assert off.numel() == self.V + 1
        w = int((off[1:] - off[:-1]).max().item()) if off.numel() > 1 else 1
        rc = self.lib.conzic_set_bert2clip(self.ctx, _ptr(off), _ptr(tok), int(tok.numel()), max(w, 1), self._stream())
        _lib.check(rc, "conzic_set_bert2clip")
        torch.cuda.current_stream(self.device).synchronize()
        self.max_tok_per_word = max(w, 1)
        self.has_table = True

    def set_text_vocab(self, tv: Dict[str, object]):
        """Uploads the device text pipeline's tables (tokens.build_text_vocab): vocabularies with '##' word pieces
        then run the whole step on the device (conzic_set_text_vocab)."""
        assert self.has_table, "call Engine.set_bert2clip first"
        keep = {k: v.to(self.device).contiguous() for k, v in tv.items() if isinstance(v, torch.Tensor)}
        t = _lib.TextVocab()
        for name in ("tok_off", "tok_bytes", "tok_cls", "tok_flags", "byte_sym", "merge_keys", "merge_vals"):
            setattr(t, name, keep[name].data_ptr())
        t.n_bytes, t.merge_bits = int(keep["tok_bytes"].numel()), int(tv["merge_bits"])
        _lib.check(self.lib.conzic_set_text_vocab(self.ctx, C.byref(t), self._stream()), "conzic_set_text_vocab")
        self.has_text_vocab = True
        self._ws = None  # the workspace plan grows by the per-candidate sequence buffer

    # ------------------------------------------------------------------ image tower (once per call)
    def set_vision(self, clip_sd: SD):
        """Uploads the CLIP vision tower (clip/clip.py:48-62) from an HF CLIPModel state dict."""
        v = "vision_model."
        n = _count_layers(clip_sd, v + "encoder.layers.{}.layer_norm1.weight")
        names = [v + "embeddings.patch_embedding.weight", v + "embeddings.class_embedding",
                 v + "embeddings.position_embedding.weight", v + "pre_layrnorm.weight", v + "pre_layrnorm.bias",
                 v + "post_layernorm.weight", v + "post_layernorm.bias", "visual_projection.weight"]
        for i in range(n):
            q = f"{v}encoder.layers.{i}."
            names += [q + "layer_norm1.weight", q + "layer_norm1.bias",
                      q + "self_attn.q_proj.weight", q + "self_attn.q_proj.bias",
                      q + "self_attn.k_proj.weight", q + "self_attn.k_proj.bias",
                      q + "self_attn.v_proj.weight", q + "self_attn.v_proj.bias",
                      q + "self_attn.out_proj.weight", q + "self_attn.out_proj.bias",
                      q + "layer_norm2.weight", q + "layer_norm2.bias",
                      q + "mlp.fc1.weight", q + "mlp.fc1.bias", q + "mlp.fc2.weight", q + "mlp.fc2.bias"]
        ts = [clip_sd[k].detach().to(self.device, torch.float32).contiguous() for k in names]
        arr = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        patch = clip_sd[names[0]]
        H, ps = patch.shape[0], patch.shape[-1]
        tokens = clip_sd[names[2]].shape[0]
        grid = int(round((tokens - 1) ** 0.5))
        vc = _lib.VisionConfig()
        vc.layers, vc.hidden, vc.heads = n, H, H // 64
        vc.ffn = clip_sd[f"{v}encoder.layers.0.mlp.fc1.weight"].shape[0]
        vc.image_size, vc.patch, vc.proj = grid * ps, ps, clip_sd["visual_projection.weight"].shape[0]
        vc.ln_eps = _dims.CLIP_LN_EPS
        _lib.check(self.lib.conzic_set_vision(self.ctx, C.byref(vc), arr, len(ts), self._stream()), "conzic_set_vision")
        del ts
        self.vision_cfg = vc
        self._vws = None

    def preprocess_uint8(self, src: torch.Tensor, cfg: dict, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """CLIPImageProcessor on the device for n same-sized images already in HBM: uint8 [n, H, W, 3] ->
        pixel_values f32[n, 3, S, S] (conzic_image_preprocess).  `cfg` = imageproc.processor_config(processor)."""
        from . import imageproc
        assert src.dtype == torch.uint8 and src.dim() == 4 and src.shape[3] == 3 and src.is_contiguous()
        m, H, W = int(src.shape[0]), int(src.shape[1]), int(src.shape[2])
        S = cfg["crop"]
        if not hasattr(self, "_img_plans"):
            self._img_plans = {}
        key = (H, W, cfg["shortest_edge"], S, cfg["mean"], cfg["std"], cfg["rescale_factor"])
        if key not in self._img_plans:
            pl = imageproc.make_plan(H, W, cfg["shortest_edge"], S, cfg["mean"], cfg["std"], cfg["rescale_factor"])
            dev = {}
            for name, ax in (("hz", pl.horiz), ("vt", pl.vert)):
                dev[name] = (torch.from_numpy(ax.weights).to(self.device), torch.from_numpy(ax.first).to(self.device),
                             torch.from_numpy(ax.count).to(self.device))
            self._img_plans[key] = (pl, dev)
        pl, dev = self._img_plans[key]
        if out is None:
            out = torch.empty((m, 3, S, S), dtype=torch.float32, device=self.device)
        ws = torch.empty(m * (pl.row_hi - pl.row_lo) * S * 3, dtype=torch.uint8, device=self.device)
        axes = []
        for name, ax in (("hz", pl.horiz), ("vt", pl.vert)):
            a = _lib.ResizeAxis()
            w, f, c = dev[name]
            a.weights, a.first, a.count = w.data_ptr(), f.data_ptr(), c.data_ptr()
            a.taps, a.precision, a.n_out, a.identity = int(ax.weights.shape[1]), int(ax.precision), S, int(ax.identity)
            axes.append(a)
        mean = (C.c_float * 3)(*pl.mean255)
        std = (C.c_float * 3)(*pl.std255)
        rc = self.lib.conzic_image_preprocess(self.ctx, _ptr(src), m, H, W, C.byref(axes[0]), C.byref(axes[1]),
                                              pl.row_lo, pl.row_hi, mean, std, _ptr(out), _ptr(ws), ws.numel(),
                                              self._stream())
        _lib.check(rc, "conzic_image_preprocess")
        return out

    def preprocess_images(self, images, cfg: dict) -> torch.Tensor:
        """The same for a list of uint8 HWC host arrays of any sizes (grouped by size, copied from pinned memory)."""
        import numpy as np
        n, S = len(images), cfg["crop"]
        out = torch.empty((n, 3, S, S), dtype=torch.float32, device=self.device)
        groups = {}
        for i, im in enumerate(images):
            groups.setdefault(im.shape[:2], []).append(i)
        for (H, W), idx in groups.items():
            host = torch.from_numpy(np.stack([images[i] for i in idx])).pin_memory()
            src = host.to(self.device, non_blocking=True)
            if len(idx) == n:
                self.preprocess_uint8(src, cfg, out)
            else:
                out[torch.tensor(idx, device=self.device)] = self.preprocess_uint8(src, cfg)
        return out

    def image_encode(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """CLIP.compute_image_representation_from_image_instance after the processor: f32[B,3,S,S] -> f32[B,proj]."""
        if getattr(self, "vision_cfg", None) is None:
            raise RuntimeError("Engine.set_vision has not been called (the checkpoint has no vision tower?)")
        pix = pixel_values.to(self.device, torch.float32).contiguous()
        B = pix.shape[0]
        assert pix.shape[1:] == (3, self.vision_cfg.image_size, self.vision_cfg.image_size), pix.shape
        out = torch.empty((B, self.vision_cfg.proj), dtype=torch.float32, device=self.device)
        need = int(self.lib.conzic_vision_workspace_bytes(self.ctx, B))
        if self._vws is None or self._vws.numel() < need:
            self._vws = None
            self._vws = torch.empty(need, dtype=torch.uint8, device=self.device)
        rc = self.lib.conzic_clip_image_encode(self.ctx, _ptr(pix), B, _ptr(out), _ptr(self._vws), self._vws.numel(),
                                               self._stream())
        _lib.check(rc, "conzic_clip_image_encode")
        return out

    # ------------------------------------------------------------------ pieces
    def bert_mlm_row(self, inp: torch.Tensor, pos: int) -> torch.Tensor:
        """logits[:, pos] of BertForMaskedLM (gen_utils.py:69,42): f32[B,V]."""
        B, L = inp.shape
        inp = inp.to(self.device, torch.int64).contiguous()
        out = torch.empty((B, self.ldl), dtype=torch.float32, device=self.device)
        ws = self.workspace(B, L, 1)
        rc = self.lib.conzic_bert_mlm_row(self.ctx, _ptr(inp), B, L, int(pos), _ptr(out), self.ldl, _ptr(ws),
                                          ws.numel(), self._stream())
        _lib.check(rc, "conzic_bert_mlm_row")
        return out[:, : self.V]

    def bert_mlm_row_padded(self, inp: torch.Tensor, pos: int) -> torch.Tensor:
        """Same as bert_mlm_row but returns the padded f32[B, ldl] buffer that gibbs_step(logits_in=...) takes."""
        B, L = inp.shape
        out = torch.empty((B, self.ldl), dtype=torch.float32, device=self.device)
        ws = self.workspace(B, L, 1)
        rc = self.lib.conzic_bert_mlm_row(self.ctx, _ptr(inp), B, L, int(pos), _ptr(out), self.ldl, _ptr(ws),
                                          ws.numel(), self._stream())
        _lib.check(rc, "conzic_bert_mlm_row")
        return out

    def topk_mask(self, logits: torch.Tensor, token_mask: torch.Tensor, temperature: float, K: int):
        """generate_caption_step (gen_utils.py:33-49) on one row of logits per image."""
        B = logits.shape[0]
        assert logits.dtype == torch.float32 and logits.stride(1) == 1
        probs = torch.empty((B, K), dtype=torch.float32, device=self.device)
        ids = torch.empty((B, K), dtype=torch.int64, device=self.device)
        tm = token_mask.reshape(-1)
        rc = self.lib.conzic_topk_mask(self.ctx, _ptr(logits), int(logits.stride(0)), B, _ptr(tm), float(temperature),
                                       int(K), _ptr(probs), _ptr(ids), self._stream())
        _lib.check(rc, "conzic_topk_mask")
        return probs, ids

    def build_clip_ids(self, inp: torch.Tensor, pos: int, ids: torch.Tensor, token_mask: torch.Tensor, T: int):
        B, L = inp.shape
        K = ids.shape[1]
        clip_ids = torch.empty((B * K, T), dtype=torch.int32, device=self.device)
        clip_len = torch.empty((B * K,), dtype=torch.int32, device=self.device)
        ids_masked = torch.empty((B, K), dtype=torch.int64, device=self.device)
        rc = self.lib.conzic_build_clip_ids(self.ctx, _ptr(inp), B, L, int(pos), _ptr(ids), _ptr(token_mask.reshape(-1)),
                                            K, _ptr(clip_ids), int(T), _ptr(clip_len), _ptr(ids_masked), self._stream())
        _lib.check(rc, "conzic_build_clip_ids")
        return clip_ids, clip_len, ids_masked

    def clip_text_encode(self, clip_ids: torch.Tensor) -> torch.Tensor:
        """CLIP.compute_text_representation after tokenisation (clip/clip.py:78-83): f32[N,512]."""
        clip_ids = clip_ids.to(self.device, torch.int32).contiguous()
        N, T = clip_ids.shape
        out = torch.empty((N, self.D), dtype=torch.float32, device=self.device)
        ws = self.workspace(N, 0, 1)
        rc = self.lib.conzic_clip_text_encode(self.ctx, _ptr(clip_ids), N, T, _ptr(out), _ptr(ws), ws.numel(),
                                              self._stream())
        _lib.check(rc, "conzic_clip_text_encode")
        return out

    def image_text_similarity(self, image_embeds: torch.Tensor, text_embeds: torch.Tensor):
        """compute_image_text_similarity_via_embeddings (clip/clip.py:86-98)."""
        B = image_embeds.shape[0]
        K = text_embeds.numel() // (B * self.D)
        score = torch.empty((B, K), dtype=torch.float32, device=self.device)
        ref = torch.empty((B, K), dtype=torch.float32, device=self.device)
        rc = self.lib.conzic_image_text_similarity(self.ctx, _ptr(text_embeds.contiguous()),
                                                   _ptr(image_embeds.contiguous()), B, K, self.logit_scale_exp,
                                                   _ptr(score), _ptr(ref), self._stream())
        _lib.check(rc, "conzic_image_text_similarity")
        return score, ref

    def score_select(self, text_embeds, image_embeds, probs, ids_masked, inp, pos, alpha, beta, gamma=None,
                     senti_raw=None, repeats=None, out_clip_ref=None, out_senti=None, out_best=None, clip_ids=None):
        """Score fuse + argmax + write-back (gen_utils.py:77-81, control_gen_utils.py:59-65) on caller-built
        candidate embeddings; `inp[:, pos]` receives the winners, `out_best` (int64[B], optional) their index.
        `clip_ids` int32[B*K, T]: the CLIP id rows `text_embeds` was encoded from -- required by the certified
        precision (candidates the bf16 scores cannot rule out are re-encoded exactly), ignored otherwise."""
        B, K = probs.shape
        T = 0
        if clip_ids is not None:
            clip_ids = clip_ids.to(self.device, torch.int32).contiguous()
            T = int(clip_ids.shape[1])
        elif self.precision == "certified":
            raise ValueError("score_select: the certified precision needs clip_ids")
        ws = self.workspace(B, inp.shape[1], K)
        if out_clip_ref is None:
            out_clip_ref = torch.empty((B,), dtype=torch.float32, device=self.device)
        ctl = gamma is not None
        if ctl and out_senti is None:
            out_senti = torch.empty((B,), dtype=torch.float32, device=self.device)
        rc = self.lib.conzic_score_select(self.ctx, _ptr(text_embeds.contiguous()), _ptr(image_embeds.contiguous()),
                                          _ptr(clip_ids), T, B, K,
                                          self.logit_scale_exp, _ptr(probs.contiguous()), _ptr(ids_masked.contiguous()),
                                          _ptr(senti_raw if ctl else None), _ptr(repeats if ctl else None),
                                          float(alpha), float(beta), float(gamma) if ctl else 0.0, _ptr(inp),
                                          inp.shape[1], int(pos), _ptr(out_clip_ref), _ptr(out_senti if ctl else None),
                                          _ptr(out_best), _ptr(ws), ws.numel(), self._stream())
        _lib.check(rc, "conzic_score_select")
        return out_clip_ref, (out_senti if ctl else None)

    def debug_linear(self, A, W, bias=None, resid=None, act: int = 0):
        M, K = A.shape
        N = W.shape[0]
        out = torch.empty((M, N), dtype=torch.float32, device=self.device)
        need = (M + N) * K * 4 + M * N * 2 + 8192
        ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        rc = self.lib.conzic_debug_linear(self.ctx, _ptr(A.contiguous()), _ptr(W.contiguous()), _ptr(bias),
                                          _ptr(resid), M, N, K, int(act), _ptr(out), _ptr(ws), ws.numel(),
                                          self._stream())
        _lib.check(rc, "conzic_debug_linear")
        return out

    # ------------------------------------------------------------------ the fused step
    def gibbs_step(self, inp: torch.Tensor, token_mask: torch.Tensor, image_embeds: torch.Tensor, pos: int,
                   dot_allowed: bool, K: int, temperature: float, alpha: float, beta: float,
                   visited_before: int, visited_after: int, gamma: Optional[float] = None,
                   senti_table: Optional[torch.Tensor] = None, out_clip_ref: Optional[torch.Tensor] = None,
                   out_senti: Optional[torch.Tensor] = None, trace: bool = False,
                   logits_in: Optional[torch.Tensor] = None):
        """One position update, in place on `inp` (int64[B,L]) and `token_mask` (f32[1,V]); gen_utils.py:66-81,
        control_gen_utils.py:45-67.  Returns (clip_ref[B], senti[B] or None, trace dict or None) -- device tensors,
        nothing is synchronised."""
        assert self.has_table, "call Engine.set_bert2clip first"
        B, L = inp.shape
        assert inp.dtype == torch.int64 and inp.is_contiguous() and inp.device == self.device
        assert token_mask.dtype == torch.float32 and token_mask.is_contiguous() and token_mask.numel() == self.V
        ctl = gamma is not None
        if out_clip_ref is None:
            out_clip_ref = torch.empty((B,), dtype=torch.float32, device=self.device)
        if ctl and out_senti is None:
            out_senti = torch.empty((B,), dtype=torch.float32, device=self.device)
        a = _lib.StepArgs()
        a.inp, a.token_mask, a.image_embeds = inp.data_ptr(), token_mask.data_ptr(), image_embeds.data_ptr()
        a.senti_table = senti_table.data_ptr() if ctl else None
        a.B, a.L, a.K, a.pos = B, L, int(K), int(pos)
        a.dot_allowed = 1 if dot_allowed else 0
        a.visited_before, a.visited_after = int(visited_before), int(visited_after)
        a.temperature, a.alpha, a.beta = float(temperature), float(alpha), float(beta)
        a.gamma = float(gamma) if ctl else 0.0
        a.logit_scale_exp = self.logit_scale_exp
        a.out_clip_ref = out_clip_ref.data_ptr()
        a.out_senti = out_senti.data_ptr() if ctl else None
        tr = None
        if trace:
            f = lambda *s: torch.empty(s, dtype=torch.float32, device=self.device)
            tr = dict(probs=f(B, K), idxs=torch.empty((B, K), dtype=torch.int64, device=self.device),
                      clip_score=f(B, K), clip_ref=f(B, K), final=f(B, K),
                      best=torch.empty((B,), dtype=torch.int64, device=self.device), logits=f(B, self.ldl))
            a.tr_probs, a.tr_ids = tr["probs"].data_ptr(), tr["idxs"].data_ptr()
            a.tr_clip_score, a.tr_clip_ref = tr["clip_score"].data_ptr(), tr["clip_ref"].data_ptr()
            a.tr_final, a.tr_best, a.tr_logits = tr["final"].data_ptr(), tr["best"].data_ptr(), tr["logits"].data_ptr()
        if logits_in is not None:  # span order: f32[B, ldl] rows from bert_mlm_row_padded
            assert logits_in.dtype == torch.float32 and logits_in.shape == (B, self.ldl) and logits_in.is_contiguous()
            a.logits_in = logits_in.data_ptr()
        ws = self.workspace(B, L, K)
        rc = self.lib.conzic_gibbs_step(self.ctx, C.byref(a), _ptr(ws), ws.numel(), self._stream())
        _lib.check(rc, "conzic_gibbs_step")
        return out_clip_ref, (out_senti if ctl else None), tr
