"""conzic_b200/imageproc.py: the plans behind the device CLIPImageProcessor (csrc/image_ops.cu).  On the CPU the
plans' integer arithmetic (imageproc.emulate, the same arithmetic the kernels run) must reproduce the real
transformers CLIPImageProcessor (torchvision backend: antialiased bicubic on uint8) -- clip/clip.py:55-58."""
import numpy as np
import pytest
import torch

from conzic_b200 import imageproc


@pytest.mark.parametrize("hw", [(300, 400), (480, 640), (100, 77), (224, 224), (231, 500), (225, 224), (640, 427), (333, 333)])
def test_plan_reproduces_the_hf_processor(hw):
    from transformers import CLIPImageProcessor
    proc = CLIPImageProcessor()
    cfg = imageproc.processor_config(proc)
    assert cfg is not None and cfg["shortest_edge"] == 224 and cfg["crop"] == 224
    H, W = hw
    img = (np.random.RandomState(H * 1000 + W).rand(H, W, 3) * 255).astype(np.uint8)
    ref = proc(images=img, return_tensors="pt")["pixel_values"][0]
    plan = imageproc.make_plan(H, W, cfg["shortest_edge"], cfg["crop"], cfg["mean"], cfg["std"], cfg["rescale_factor"])
    got = imageproc.emulate(img, plan)
    assert got.shape == ref.shape == (3, 224, 224)
    assert float((got - ref).abs().max()) <= 1e-6
    assert plan.row_hi - plan.row_lo <= H and plan.horiz.weights.dtype == np.int16


def test_inputs_and_unsupported_processors():
    from PIL import Image
    arr = (np.random.RandomState(0).rand(50, 60, 3) * 255).astype(np.uint8)
    assert np.array_equal(imageproc.to_uint8_hwc(Image.fromarray(arr)), arr)
    assert np.array_equal(imageproc.to_uint8_hwc(arr.transpose(2, 0, 1)), arr)
    assert imageproc.to_uint8_hwc(torch.zeros(3, 4, 4)) is None
    assert imageproc.processor_config(object()) is None


@pytest.mark.gpu
def test_device_image_preprocess_matches_hf_processor():
    """conzic_image_preprocess (csrc/image_ops.cu) against the real transformers CLIPImageProcessor on PIL images of
    mixed sizes (down- and up-scaling, an axis that is not resized, odd crops): pixel values within 1e-6, and the
    CLIP wrapper's image route (pre-process + image tower on the device) equal to pre-processing on the host."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import gpu_common as gc
    from PIL import Image
    from transformers import CLIPImageProcessor
    from conzic_b200 import runtime
    from conzic_b200.clip.clip import CLIP
    from synthetic import synth
    proc = CLIPImageProcessor()
    sizes = [(300, 400), (480, 640), (100, 77), (224, 224), (231, 500), (640, 427), (300, 400), (225, 224)]
    rs = np.random.RandomState(3)
    imgs = [Image.fromarray((rs.rand(h, w, 3) * 255).astype(np.uint8)) for h, w in sizes]
    ref = proc(images=imgs, return_tensors="pt")["pixel_values"]
    runtime.clear()
    clip = CLIP(state_dict=synth.make_clip_state_dict(0, vision=True), tokenizer=synth.SynthCLIPTokenizer(),
                processor=proc).to("cuda")
    eng = clip._engine()
    cfg = imageproc.processor_config(proc)
    got = eng.preprocess_images([imageproc.to_uint8_hwc(im) for im in imgs], cfg)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    assert float((got.cpu() - ref).abs().max()) <= 1e-6
    emb = clip.compute_image_representation_from_image_instance(imgs)
    emb_ref = clip.compute_image_representation_from_pixels(ref)
    assert float((emb - emb_ref).abs().max()) <= 1e-5 * float(emb_ref.abs().max())
    runtime.clear()
