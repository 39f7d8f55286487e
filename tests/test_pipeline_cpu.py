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


def _fake_ops(monkeypatch):
    """torch definitions of the three device ops the loop calls (rf_axpby_f16, rf_cfg_pndm_step_f16, device slerp off)"""
    from riffusion import tc_ops

    def axpby(x, noise, a, b, mask=None, z=None):
        y = a * x.float() + b * noise.float()
        if mask is not None:
            y = y * mask.float() + z.float() * (1 - mask.float())
        return y.to(x.dtype)

    def cfg_step(eps_pair, guidance, hist, coef, sample, ca, cb, want_eps=True):
        n = sample.shape[0]
        eu, et = eps_pair[:n].float(), eps_pair[n:].float()
        eps = eu + guidance * (et - eu)
        e = coef[0] * eps
        for c, h in zip(coef[1:], hist):
            e = e + c * h.float()
        return (eps.to(sample.dtype) if want_eps else None), (ca * sample.float() - cb * e).to(sample.dtype)

    monkeypatch.setattr(tc_ops, "axpby", axpby)
    monkeypatch.setattr(tc_ops, "cfg_pndm_step", cfg_step)


def test_interpolate_img2img_control_flow_matches_oracle_loop(monkeypatch):
    """strength interpolation, init_timestep / t_start arithmetic, noise slerp, add_noise, CFG doubling, PLMS stepping and
    the mask blend of interpolate_img2img (reference :311-425) against the oracle loop, with a smooth stand-in for the
    UNet so that the whole thing runs on the CPU (fp16 tensors like the product, fp32 oracle)"""
    from oracle import unet_oracle as uo

    _fake_ops(monkeypatch)

    def model(x, t, ctx):
        return 0.3 * torch.tanh(x.float()) + 0.002 * (t / 1000.0) + 0.05 * ctx.float().mean(dim=(1, 2))[:, None, None, None]

    class FakeUNet:
        def __call__(self, x, t, encoder_hidden_states=None, **kw):
            return types.SimpleNamespace(sample=model(x, int(t), encoder_hidden_states).to(torch.float16))

    pipe = RiffusionPipeline(vae=None, unet=FakeUNet(), device="cpu")
    pipe.use_cuda_graph = False
    pipe.device_slerp = False                      # the reference's host-numpy slerp
    torch.manual_seed(5)
    lat = torch.randn(1, 4, 8, 8).half()
    na, nb = torch.randn(1, 4, 8, 8).half(), torch.randn(1, 4, 8, 8).half()
    text, uncond = torch.randn(1, 77, 16).half(), torch.randn(1, 77, 16).half()
    mask = (torch.rand(1, 4, 8, 8) > 0.5).half()
    for steps, sa, sb, alpha, m in ((50, 0.75, 0.75, 0.5, None), (20, 0.5, 0.9, 0.25, None), (10, 1.0, 1.0, 0.0, mask)):
        strength = (1 - alpha) * sa + alpha * sb
        ref, n_ref = uo.img2img_loop(lambda x, t, c: model(x, t, c), uo.PNDMSchedulerOracle(), text.float(), uncond.float(),
                                     lat.float(), na.float(), nb.float(), alpha, strength, steps, 7.0,
                                     mask=None if m is None else m.float())
        out = pipe.interpolate_img2img(text_embeddings=text, init_latents=lat, generator_a=None, generator_b=None,
                                       interpolate_alpha=alpha, strength_a=sa, strength_b=sb, num_inference_steps=steps,
                                       guidance_scale=7.0, uncond_embeddings=uncond, noise_a=na, noise_b=nb, mask=m,
                                       output_type="latent")
        assert out["n_unet_evals"] == n_ref, (steps, sa, sb)
        err = float((out["latents_unscaled"].float() - ref).norm() / ref.norm())
        assert err < 2e-2, (steps, err)            # fp16 latents through up to 38 guided steps


def test_datatypes_schema_and_from_dict():
    """same fields / order / defaults as riffusion/datatypes.py:10-73; from_dict = the server's dacite construction"""
    import dataclasses

    import pytest

    from riffusion.datatypes import InferenceOutput

    assert [f.name for f in dataclasses.fields(PromptInput)] == ["prompt", "seed", "negative_prompt", "denoising", "guidance"]
    assert [f.name for f in dataclasses.fields(InferenceInput)] == ["start", "end", "alpha", "num_inference_steps",
                                                                    "seed_image_id", "mask_image_id"]
    assert [f.name for f in dataclasses.fields(InferenceOutput)] == ["image", "audio", "duration_s"]
    p = PromptInput("a", 1)
    assert (p.negative_prompt, p.denoising, p.guidance) == (None, 0.75, 7.0)
    req = InferenceInput.from_dict({"alpha": 0.75, "num_inference_steps": 50, "seed_image_id": "og_beat",
                                    "start": {"prompt": "church bells on sunday", "seed": 42},
                                    "end": {"prompt": "jazz with piano", "seed": 123, "denoising": 0.8}})     # README.md:155-171
    assert req.start == PromptInput("church bells on sunday", 42) and req.end.denoising == 0.8 and req.mask_image_id is None
    with pytest.raises(KeyError):
        InferenceInput.from_dict({"alpha": 0.1, "start": {"prompt": "x", "seed": 1, "bogus": 2}, "end": {"prompt": "y", "seed": 2}})
    with pytest.raises(dataclasses.FrozenInstanceError):
        req.alpha = 0.5
