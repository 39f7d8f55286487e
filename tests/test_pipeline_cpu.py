"""Host control flow of RiffusionPipeline.riffuse (riffusion/riffusion_pipeline.py:208-287 of the reference) with the
device work replaced by recorders: guidance / prompt interpolation, generator seeds, mask preparation and the argument
mapping into interpolate_img2img must match the reference line by line.  No GPU, no kernels."""
import types

import numpy as np
import torch
from PIL import Image

from riffusion.datatypes import InferenceInput, PromptInput
from riffusion.riffusion_pipeline import RiffusionPipeline, preprocess_image, preprocess_mask


def _pipe():
    p = RiffusionPipeline(vae=types.SimpleNamespace(config=types.SimpleNamespace(block_out_channels=[128, 256, 512, 512])),
                          unet=None, device="cpu")
    emb = {"church bells on sunday": torch.full((1, 77, 8), 1.0), "jazz with piano": torch.full((1, 77, 8), 3.0)}
    calls = {}
    p.embed_text_weighted = lambda text: emb[text]
    p.embed_text = lambda text: emb[text] + 100

    def encode_image(img, generator):
        calls["encode"] = (img.size, generator.initial_seed())
        return torch.zeros(1, 4, 64, 64)

    def interpolate(**kw):
        calls["loop"] = kw
        return {"images": ["IMAGE0", "IMAGE1"]}

    p.encode_image = encode_image
    p.interpolate_img2img = interpolate
    return p, calls


def test_riffuse_argument_mapping():
    p, calls = _pipe()
    inputs = InferenceInput(alpha=0.25, num_inference_steps=37, seed_image_id="og_beat",
                            start=PromptInput("church bells on sunday", seed=42, denoising=0.6, guidance=6.0),
                            end=PromptInput("jazz with piano", seed=123, denoising=0.9, guidance=8.0))
    img = Image.new("RGB", (512, 512), (10, 20, 30))
    out = p.riffuse(inputs, init_image=img)
    assert out == "IMAGE0"                                              # outputs["images"][0]  (:287)
    kw = calls["loop"]
    assert abs(kw["guidance_scale"] - (6.0 * 0.75 + 8.0 * 0.25)) < 1e-12  # :231
    assert torch.equal(kw["text_embeddings"], torch.full((1, 77, 8), 1.0 + 0.25 * 2.0))   # linear, not slerp (:249)
    assert kw["generator_a"].initial_seed() == 42 and kw["generator_b"].initial_seed() == 123   # :238-239
    assert calls["encode"] == ((512, 512), 42)                           # posterior noise from start.seed (:259-263)
    assert kw["interpolate_alpha"] == 0.25 and kw["strength_a"] == 0.6 and kw["strength_b"] == 0.9
    assert kw["num_inference_steps"] == 37 and kw["mask"] is None
    # use_reweighting=False routes through embed_text (:241-246)
    p.riffuse(inputs, init_image=img, use_reweighting=False)
    assert torch.equal(calls["loop"]["text_embeddings"], torch.full((1, 77, 8), 101.0 + 0.25 * 2.0))


def test_riffuse_mask_and_preprocess():
    p, calls = _pipe()
    inputs = InferenceInput(alpha=0.0, start=PromptInput("church bells on sunday", seed=1),
                            end=PromptInput("jazz with piano", seed=2))
    mask_img = Image.fromarray((np.arange(520 * 530).reshape(520, 530) % 256).astype(np.uint8))
    p.riffuse(inputs, init_image=Image.new("RGB", (530, 520)), mask_image=mask_img)
    m = calls["loop"]["mask"]
    assert m.shape == (1, 4, 512 // 8, 512 // 8)                         # sizes rounded down to multiples of 32, /8 (:455-477)
    ref = np.array(mask_img.convert("L").resize((64, 64), resample=Image.NEAREST)).astype(np.float32) / 255.0
    assert np.allclose(m[0, 2].numpy(), 1 - ref)                         # white = repaint, tiled over the 4 latent channels
    x = preprocess_image(Image.new("RGB", (530, 520), (255, 0, 127)))
    assert x.shape == (1, 3, 512, 512) and float(x[0, 0].min()) == 1.0 and float(x[0, 1].max()) == -1.0
    assert abs(float(x[0, 2, 0, 0]) - (2 * 127 / 255 - 1)) < 1e-6
    assert preprocess_mask(mask_img, scale_factor=8).dtype == torch.float32
