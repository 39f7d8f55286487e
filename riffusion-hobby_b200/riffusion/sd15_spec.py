"""Parameter names and shapes of the SD-1.5 UNet2DConditionModel and AutoencoderKL in diffusers' state_dict
layout, generated from the architecture description (block_out_channels, layers_per_block, ...).

Used to create random-init weights (BASELINE config 4: "random-init SD-1.5 UNet weights" — there is no network
to fetch riffusion/riffusion-model-v1) and to validate checkpoints before repacking them for the kernels.
"""
from __future__ import annotations

import typing as T

import torch

Spec = T.List[T.Tuple[str, T.Tuple[int, ...]]]


def _conv(name: str, cout: int, cin: int, k: int) -> Spec:
    return [(name + ".weight", (cout, cin, k, k)), (name + ".bias", (cout,))]


def _lin(name: str, cout: int, cin: int, bias: bool = True) -> Spec:
    return [(name + ".weight", (cout, cin))] + ([(name + ".bias", (cout,))] if bias else [])


def _norm(name: str, c: int) -> Spec:
    return [(name + ".weight", (c,)), (name + ".bias", (c,))]


def _resnet(p: str, cin: int, cout: int, temb: T.Optional[int]) -> Spec:
    s = _norm(p + "norm1", cin) + _conv(p + "conv1", cout, cin, 3)
    if temb:
        s += _lin(p + "time_emb_proj", cout, temb)
    s += _norm(p + "norm2", cout) + _conv(p + "conv2", cout, cout, 3)
    if cin != cout:
        s += _conv(p + "conv_shortcut", cout, cin, 1)
    return s


def _transformer(p: str, c: int, ctx: int) -> Spec:
    s = _norm(p + "norm", c) + _conv(p + "proj_in", c, c, 1)
    t = p + "transformer_blocks.0."
    for attn, kdim in (("attn1", c), ("attn2", ctx)):
        s += _lin(t + attn + ".to_q", c, c, False) + _lin(t + attn + ".to_k", c, kdim, False)
        s += _lin(t + attn + ".to_v", c, kdim, False) + _lin(t + attn + ".to_out.0", c, c)
    s += _lin(t + "ff.net.0.proj", 8 * c, c) + _lin(t + "ff.net.2", c, 4 * c)
    for n in ("norm1", "norm2", "norm3"):
        s += _norm(t + n, c)
    return s + _conv(p + "proj_out", c, c, 1)


def unet_spec(block_out_channels=(320, 640, 1280, 1280), cross_attention_dim=768, in_channels=4, out_channels=4) -> Spec:
    c = list(block_out_channels)
    temb = 4 * c[0]
    s = _conv("conv_in", c[0], in_channels, 3)
    s += _lin("time_embedding.linear_1", temb, c[0]) + _lin("time_embedding.linear_2", temb, temb)
    skips = [c[0]]
    prev = c[0]
    for i, co in enumerate(c):
        last = i == len(c) - 1
        for j in range(2):
            s += _resnet(f"down_blocks.{i}.resnets.{j}.", prev if j == 0 else co, co, temb)
            if not last:
                s += _transformer(f"down_blocks.{i}.attentions.{j}.", co, cross_attention_dim)
            skips.append(co)
        if not last:
            s += _conv(f"down_blocks.{i}.downsamplers.0.conv", co, co, 3)
            skips.append(co)
        prev = co
    s += _resnet("mid_block.resnets.0.", c[-1], c[-1], temb) + _transformer("mid_block.attentions.0.", c[-1], cross_attention_dim)
    s += _resnet("mid_block.resnets.1.", c[-1], c[-1], temb)
    for i, co in enumerate(c[::-1]):
        for j in range(3):
            s += _resnet(f"up_blocks.{i}.resnets.{j}.", (prev if j == 0 else co) + skips.pop(), co, temb)
            if i > 0:
                s += _transformer(f"up_blocks.{i}.attentions.{j}.", co, cross_attention_dim)
        if i < len(c) - 1:
            s += _conv(f"up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        prev = co
    return s + _norm("conv_norm_out", c[0]) + _conv("conv_out", out_channels, c[0], 3)


def _vae_mid(p: str, c: int) -> Spec:
    s = _resnet(p + "resnets.0.", c, c, None)
    a = p + "attentions.0."
    s += _norm(a + "group_norm", c)
    for n in ("query", "key", "value", "proj_attn"):
        s += _lin(a + n, c, c)
    return s + _resnet(p + "resnets.1.", c, c, None)


def vae_spec(block_out_channels=(128, 256, 512, 512), latent_channels=4) -> Spec:
    c = list(block_out_channels)
    s = _conv("encoder.conv_in", c[0], 3, 3)
    prev = c[0]
    for i, co in enumerate(c):
        for j in range(2):
            s += _resnet(f"encoder.down_blocks.{i}.resnets.{j}.", prev if j == 0 else co, co, None)
        if i < len(c) - 1:
            s += _conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        prev = co
    s += _vae_mid("encoder.mid_block.", c[-1]) + _norm("encoder.conv_norm_out", c[-1])
    s += _conv("encoder.conv_out", 2 * latent_channels, c[-1], 3)
    rev = c[::-1]
    s += _conv("decoder.conv_in", rev[0], latent_channels, 3) + _vae_mid("decoder.mid_block.", rev[0])
    prev = rev[0]
    for i, co in enumerate(rev):
        for j in range(3):
            s += _resnet(f"decoder.up_blocks.{i}.resnets.{j}.", prev if j == 0 else co, co, None)
        if i < len(c) - 1:
            s += _conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        prev = co
    s += _norm("decoder.conv_norm_out", c[0]) + _conv("decoder.conv_out", 3, c[0], 3)
    s += _conv("quant_conv", 2 * latent_channels, 2 * latent_channels, 1)
    return s + _conv("post_quant_conv", latent_channels, latent_channels, 1)


def random_state_dict(spec: Spec, seed: int, std: float = 0.02, dtype=torch.float16) -> T.Dict[str, torch.Tensor]:
    """N(0, std^2) matrices / kernels, norm scales 1, biases 0 (BASELINE config 4)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in spec:
        if len(shape) > 1:
            out[name] = (torch.randn(shape, generator=g) * std).to(dtype)
        elif name.endswith(".weight"):
            out[name] = torch.ones(shape, dtype=dtype)
        else:
            out[name] = torch.zeros(shape, dtype=dtype)
    return out


def random_state_dicts(seed: int = 0, with_vae: bool = True):
    unet = random_state_dict(unet_spec(), seed)
    vae = random_state_dict(vae_spec(), seed + 1) if with_vae else None
    return unet, vae
