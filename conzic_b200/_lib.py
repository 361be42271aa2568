"""ctypes binding of libconzic.so (the C ABI in include/conzic.h).

The library is the product: if it cannot be loaded (not built, or no nvcc to build it) importing the
engine fails loudly.  There is no CPU or PyTorch fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libconzic.so")

PREC_BF16, PREC_BF16X3, PREC_CERTIFIED = 0, 1, 2
FLAG_NO_PDL, FLAG_LN_STANDALONE, FLAG_WIDE_LSU, FLAG_LSU_OUT = 1, 2, 4, 16
CERT_STATS = 8
GEMM_TCGEN05, GEMM_SIMT_DEBUG = 0, 1
BERT_GLOBALS, CLIP_GLOBALS, PER_LAYER = 10, 5, 16

# every symbol include/conzic.h declares; tests check the shared object exports all of them
EXPORTS = [
    "conzic_abi_version", "conzic_last_error", "conzic_ctx_create", "conzic_ctx_destroy", "conzic_set_bert2clip",
    "conzic_workspace_bytes", "conzic_bert_mlm_row", "conzic_topk_mask", "conzic_build_clip_ids",
    "conzic_clip_text_encode", "conzic_image_text_similarity", "conzic_gibbs_step", "conzic_launch_count",
    "conzic_debug_linear", "conzic_profile", "conzic_profile_read", "conzic_cert_stats",
    "conzic_set_vision", "conzic_vision_workspace_bytes", "conzic_clip_image_encode", "conzic_score_select", "conzic_set_text_vocab", "conzic_image_preprocess",
]


class Config(C.Structure):
    _fields_ = [
        ("bert_layers", C.c_int32), ("bert_hidden", C.c_int32), ("bert_heads", C.c_int32), ("bert_ffn", C.c_int32),
        ("bert_vocab", C.c_int32), ("bert_maxpos", C.c_int32), ("bert_ln_eps", C.c_float),
        ("clip_layers", C.c_int32), ("clip_hidden", C.c_int32), ("clip_heads", C.c_int32), ("clip_ffn", C.c_int32),
        ("clip_vocab", C.c_int32), ("clip_maxpos", C.c_int32), ("clip_proj", C.c_int32), ("clip_ln_eps", C.c_float),
        ("pad_id", C.c_int32), ("unk_id", C.c_int32), ("cls_id", C.c_int32), ("sep_id", C.c_int32),
        ("mask_id", C.c_int32), ("dot_id", C.c_int32), ("clip_bos", C.c_int32), ("clip_eos", C.c_int32),
        ("precision", C.c_int32), ("gemm_impl", C.c_int32), ("clip_chunk_rows", C.c_int32),
        ("cert_dcos", C.c_float), ("cert_dcos_lo", C.c_float), ("cert_zratio_lo", C.c_float), ("cert_zratio_hi", C.c_float),
        ("cert_fcap", C.c_int32), ("flags", C.c_int32),
    ]


class TextVocab(C.Structure):
    _fields_ = [("tok_off", C.c_void_p), ("tok_bytes", C.c_void_p), ("tok_cls", C.c_void_p), ("tok_flags", C.c_void_p),
                ("byte_sym", C.c_void_p), ("merge_keys", C.c_void_p), ("merge_vals", C.c_void_p),
                ("n_bytes", C.c_int32), ("merge_bits", C.c_int32)]


class ResizeAxis(C.Structure):
    _fields_ = [("weights", C.c_void_p), ("first", C.c_void_p), ("count", C.c_void_p), ("taps", C.c_int32),
                ("precision", C.c_int32), ("n_out", C.c_int32), ("identity", C.c_int32)]


class VisionConfig(C.Structure):
    _fields_ = [("layers", C.c_int32), ("hidden", C.c_int32), ("heads", C.c_int32), ("ffn", C.c_int32),
                ("image_size", C.c_int32), ("patch", C.c_int32), ("proj", C.c_int32), ("ln_eps", C.c_float)]


class StepArgs(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("token_mask", C.c_void_p), ("image_embeds", C.c_void_p), ("senti_table", C.c_void_p),
        ("B", C.c_int32), ("L", C.c_int32), ("K", C.c_int32), ("pos", C.c_int32),
        ("dot_allowed", C.c_int32), ("visited_before", C.c_int32), ("visited_after", C.c_int32),
        ("temperature", C.c_float), ("alpha", C.c_float), ("beta", C.c_float), ("gamma", C.c_float),
        ("logit_scale_exp", C.c_float),
        ("out_clip_ref", C.c_void_p), ("out_senti", C.c_void_p),
        ("tr_probs", C.c_void_p), ("tr_ids", C.c_void_p), ("tr_clip_score", C.c_void_p), ("tr_clip_ref", C.c_void_p),
        ("tr_final", C.c_void_p), ("tr_best", C.c_void_p), ("tr_logits", C.c_void_p),
        ("logits_in", C.c_void_p),
    ]


_lib = None


def _declare(lib):
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    lib.conzic_abi_version.restype = C.c_int
    lib.conzic_last_error.restype = C.c_char_p
    lib.conzic_ctx_create.restype = C.c_int
    lib.conzic_ctx_create.argtypes = [C.POINTER(Config), C.POINTER(vp), i32, C.POINTER(vp), i32, vp, C.POINTER(vp)]
    lib.conzic_ctx_destroy.restype = None
    lib.conzic_ctx_destroy.argtypes = [vp]
    lib.conzic_set_bert2clip.restype = C.c_int
    lib.conzic_set_bert2clip.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.conzic_workspace_bytes.restype = sz
    lib.conzic_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.conzic_bert_mlm_row.restype = C.c_int
    lib.conzic_bert_mlm_row.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp, sz, vp]
    lib.conzic_topk_mask.restype = C.c_int
    lib.conzic_topk_mask.argtypes = [vp, vp, i32, i32, vp, f32, i32, vp, vp, vp]
    lib.conzic_build_clip_ids.restype = C.c_int
    lib.conzic_build_clip_ids.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, vp, i32, vp, vp, vp]
    lib.conzic_clip_text_encode.restype = C.c_int
    lib.conzic_clip_text_encode.argtypes = [vp, vp, i32, i32, vp, vp, sz, vp]
    lib.conzic_image_text_similarity.restype = C.c_int
    lib.conzic_image_text_similarity.argtypes = [vp, vp, vp, i32, i32, f32, vp, vp, vp]
    lib.conzic_set_text_vocab.restype = C.c_int
    lib.conzic_set_text_vocab.argtypes = [vp, C.POINTER(TextVocab), vp]
    lib.conzic_score_select.restype = C.c_int
    lib.conzic_score_select.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, vp, vp, vp, vp, f32, f32, f32, vp, i32, i32,
                                        vp, vp, vp, vp, sz, vp]
    lib.conzic_gibbs_step.restype = C.c_int
    lib.conzic_gibbs_step.argtypes = [vp, C.POINTER(StepArgs), vp, sz, vp]
    lib.conzic_launch_count.restype = C.c_uint64
    lib.conzic_launch_count.argtypes = [vp]
    lib.conzic_debug_linear.restype = C.c_int
    lib.conzic_debug_linear.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, sz, vp]
    lib.conzic_cert_stats.restype = C.c_int
    lib.conzic_cert_stats.argtypes = [vp, C.POINTER(C.c_uint64), i32]
    lib.conzic_set_vision.restype = C.c_int
    lib.conzic_set_vision.argtypes = [vp, C.POINTER(VisionConfig), C.POINTER(vp), i32, vp]
    lib.conzic_vision_workspace_bytes.restype = sz
    lib.conzic_vision_workspace_bytes.argtypes = [vp, i32]
    lib.conzic_clip_image_encode.restype = C.c_int
    lib.conzic_clip_image_encode.argtypes = [vp, vp, i32, vp, vp, sz, vp]
    lib.conzic_image_preprocess.restype = C.c_int
    lib.conzic_image_preprocess.argtypes = [vp, vp, i32, i32, i32, C.POINTER(ResizeAxis), C.POINTER(ResizeAxis), i32, i32,
                                            C.POINTER(C.c_float), C.POINTER(C.c_float), vp, vp, sz, vp]
    lib.conzic_profile.restype = C.c_int
    lib.conzic_profile.argtypes = [vp, i32]
    lib.conzic_profile_read.restype = C.c_int
    lib.conzic_profile_read.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    return lib


def load(build_if_missing: bool = True):
    """Returns the loaded library; raises if it is absent and cannot be built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing; run `python -m conzic_b200.build`")
        from . import build as _build
        _build.build()
    _lib = _declare(C.CDLL(LIB_PATH))
    if _lib.conzic_abi_version() != 6:
        raise RuntimeError("libconzic.so ABI version mismatch; rebuild with `python -m conzic_b200.build --force`")
    return _lib


def last_error() -> str:
    return (load().conzic_last_error() or b"").decode()


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")
