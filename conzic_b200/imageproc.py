"""CLIPImageProcessor on the device: plans for csrc/image_ops.cu (conzic_image_preprocess).

The reference turns PIL images into pixel tensors with `self.processor(images=image, return_tensors="pt")`
(clip/clip.py:55-58).  With transformers 5.x that is the torchvision backend: uint8 CHW tensor -> resize of the
shortest edge to 224 with ANTIALIASED BICUBIC interpolation on uint8 (ATen's fixed-point kernel: a horizontal pass,
then a vertical pass, int16 filter taps, rounding to uint8 after each pass) -> centre crop 224 x 224 -> one fused
`(x - 255 mean) / (255 std)` in fp32.  This module restates the host-side arithmetic of that kernel -- which input
pixels each output pixel reads and the integer taps -- so that the device kernels reproduce the processor's output
(tests/test_imageproc.py: the emulation is bit-exact against torch on the CPU; -m gpu: the kernels against the
real CLIPImageProcessor).  Only the output rows / columns inside the centre crop are planned and computed.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _cubic(x: np.ndarray) -> np.ndarray:
    """Keys' bicubic kernel with a = -0.5 (ATen aa_filter for bicubic, taken from Pillow)."""
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0,
                    np.where(x < 2.0, (((x - 5.0) * x + 8.0) * x - 4.0) * a, 0.0))


@dataclass
class AxisPlan:
    weights: np.ndarray  # int16 [n_out, taps]
    first: np.ndarray    # int32 [n_out]
    count: np.ndarray    # int32 [n_out]
    precision: int
    identity: bool       # in == out: no pass along this axis (the kernel just crops)


def axis_plan(in_size: int, out_size: int, crop_lo: int, crop_n: int) -> AxisPlan:
    """Filter taps of output positions [crop_lo, crop_lo + crop_n) when `in_size` samples are resized to `out_size`
    (ATen UpSampleKernel: _compute_indices_min_size_weights_aa + the int16 conversion of the uint8 path)."""
    if in_size == out_size:
        first = np.arange(crop_lo, crop_lo + crop_n, dtype=np.int32)
        return AxisPlan(np.ones((crop_n, 1), np.int16), first, np.ones(crop_n, np.int32), 0, True)
    scale = in_size / out_size
    support = 2.0 * scale if scale >= 1.0 else 2.0
    taps = int(math.ceil(support)) * 2 + 1
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    W = np.zeros((out_size, taps), np.float64)
    first = np.zeros(out_size, np.int64)
    count = np.zeros(out_size, np.int64)
    for i in range(out_size):  # every output position: the int16 precision depends on the largest tap of the whole axis
        center = scale * (i + 0.5)
        lo = max(int(center - support + 0.5), 0)
        n = min(int(center + support + 0.5), in_size) - lo
        w = _cubic((np.arange(n) + lo - center + 0.5) * invscale)
        tot = 0.0
        for v in w:  # the kernel sums in this order
            tot += v
        if tot != 0.0:
            w = w / tot
        W[i, :n] = w
        first[i], count[i] = lo, n
    wmax = float(W.max())
    prec = 0
    for prec in range(0, 22):
        if int(0.5 + wmax * (1 << (prec + 1))) >= (1 << 15):
            break
    Wi = np.trunc(np.where(W < 0, -0.5 + W * (1 << prec), 0.5 + W * (1 << prec))).astype(np.int16)
    sl = slice(crop_lo, crop_lo + crop_n)
    return AxisPlan(np.ascontiguousarray(Wi[sl]), first[sl].astype(np.int32), count[sl].astype(np.int32), prec, False)


def resized_size(h: int, w: int, shortest_edge: int) -> Tuple[int, int]:
    """HF get_resize_output_image_size(size={'shortest_edge': s}, default_to_square=False)."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = shortest_edge, int(shortest_edge * long_ / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


@dataclass
class ImagePlan:
    H: int
    W: int
    out: int
    horiz: AxisPlan
    vert: AxisPlan
    row_lo: int   # input rows the vertical pass reads: [row_lo, row_hi)
    row_hi: int
    mean255: Tuple[float, float, float]
    std255: Tuple[float, float, float]


def make_plan(H: int, W: int, shortest_edge: int = 224, crop: int = 224, mean: Sequence[float] = CLIP_MEAN,
              std: Sequence[float] = CLIP_STD, rescale_factor: float = 1.0 / 255.0) -> ImagePlan:
    nh, nw = resized_size(H, W, shortest_edge)
    if nh < crop or nw < crop:
        raise ValueError("the resized image is smaller than the crop")
    top = int((nh - crop) / 2.0)   # HF TorchvisionBackend.center_crop
    left = int((nw - crop) / 2.0)
    horiz = axis_plan(W, nw, left, crop)
    vert = axis_plan(H, nh, top, crop)
    row_lo = int(vert.first.min())
    row_hi = int((vert.first + vert.count).max())
    m = (torch.tensor(list(mean)) * (1.0 / rescale_factor)).tolist()  # the processor's fused rescale + normalise, fp32
    s = (torch.tensor(list(std)) * (1.0 / rescale_factor)).tolist()
    return ImagePlan(H, W, crop, horiz, vert, row_lo, row_hi, tuple(m), tuple(s))


def emulate(img: np.ndarray, plan: ImagePlan) -> torch.Tensor:
    """The device kernels' arithmetic in numpy (CPU check of the plans): uint8 HWC -> f32 [3, out, out]."""
    x = img.astype(np.int64)
    hp, vp = plan.horiz, plan.vert
    rows = x[plan.row_lo: plan.row_hi]
    if hp.identity:
        tmp = rows[:, hp.first]
    else:
        tmp = np.zeros((rows.shape[0], plan.out, 3), np.int64)
        for i in range(plan.out):
            acc = np.full((rows.shape[0], 3), 1 << (hp.precision - 1), np.int64)
            for j in range(int(hp.count[i])):
                acc += rows[:, hp.first[i] + j] * int(hp.weights[i, j])
            tmp[:, i] = np.clip(acc >> hp.precision, 0, 255)
    if vp.identity:
        out = tmp[vp.first - plan.row_lo]
    else:
        out = np.zeros((plan.out, plan.out, 3), np.int64)
        for i in range(plan.out):
            acc = np.full((plan.out, 3), 1 << (vp.precision - 1), np.int64)
            for j in range(int(vp.count[i])):
                acc += tmp[vp.first[i] - plan.row_lo + j] * int(vp.weights[i, j])
            out[i] = np.clip(acc >> vp.precision, 0, 255)
    t = torch.from_numpy(out.astype(np.float32)).permute(2, 0, 1)
    m = torch.tensor(plan.mean255, dtype=torch.float32).view(3, 1, 1)
    s = torch.tensor(plan.std255, dtype=torch.float32).view(3, 1, 1)
    return (t - m) / s


def processor_config(processor) -> Optional[dict]:
    """The settings of a Hugging Face CLIPProcessor / CLIPImageProcessor if they are the ones the device path
    restates (shortest-edge bicubic resize, centre crop to a square, rescale, normalise); None otherwise."""
    ip = getattr(processor, "image_processor", processor)
    try:
        d = ip.to_dict()
    except Exception:  # noqa: BLE001
        return None
    size, crop = d.get("size") or {}, d.get("crop_size") or {}
    ok = (d.get("do_resize") and d.get("do_center_crop") and d.get("do_rescale") and d.get("do_normalize")
          and "shortest_edge" in size and crop.get("height") == crop.get("width") and crop.get("height") is not None
          and int(d.get("resample", -1)) == 3)  # PIL.Image.BICUBIC
    if not ok:
        return None
    backend = type(ip).__mro__[1].__name__ if len(type(ip).__mro__) > 1 else ""
    if "Torchvision" not in backend and "Fast" not in type(ip).__name__:
        return None  # the PIL backend resizes with Pillow's own fixed-point kernel: not restated here
    return dict(shortest_edge=int(size["shortest_edge"]), crop=int(crop["height"]), mean=tuple(d["image_mean"]),
                std=tuple(d["image_std"]), rescale_factor=float(d["rescale_factor"]),
                convert_rgb=bool(d.get("do_convert_rgb", True)))


def to_uint8_hwc(image) -> Optional[np.ndarray]:
    """PIL image / uint8 HWC or CHW array -> contiguous uint8 [H, W, 3]; None if the input is something else."""
    if hasattr(image, "convert") and hasattr(image, "size"):  # PIL
        return np.ascontiguousarray(np.asarray(image.convert("RGB"), dtype=np.uint8))
    if isinstance(image, torch.Tensor):
        image = image.detach().cpu().numpy()
    if isinstance(image, np.ndarray) and image.dtype == np.uint8 and image.ndim == 3:
        if image.shape[2] == 3:
            return np.ascontiguousarray(image)
        if image.shape[0] == 3:
            return np.ascontiguousarray(image.transpose(1, 2, 0))
    return None
