"""Oracle for the VAE of hot path (b): plain-PyTorch restatement of diffusers 0.9 `AutoencoderKL`
(SD-1.x config: block_out_channels (128,256,512,512), layers_per_block 2, latent_channels 4, norm_num_groups 32,
eps 1e-6) as reached from riffusion/riffusion_pipeline.py:255-264 (`vae.encode(...).latent_dist.sample`) and
:427-428 (`vae.decode(latents / 0.18215).sample`).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (diffusers absent; restated from memory).  Self-check: parameter count
83,653,863 (the published 83.7 M of the SD VAE).  Parameter names follow the diffusers state_dict.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle.unet_oracle import ResnetBlock2D


class AttentionBlock(nn.Module):
    """single-head spatial self-attention of the VAE mid block (diffusers AttentionBlock)"""

    def __init__(self, c: int, groups: int = 32):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.query, self.key, self.value, self.proj_attn = (nn.Linear(c, c) for _ in range(4))

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.query(h), self.key(h), self.value(h)
        attn = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * C ** -0.5, dim=-1)
        h = self.proj_attn(torch.bmm(attn, v)).transpose(1, 2).reshape(B, C, H, W)
        return h + x


class VaeMid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, groups, 1e-6), ResnetBlock2D(c, c, None, groups, 1e-6)])
        self.attentions = nn.ModuleList([AttentionBlock(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class VaeUp(nn.Module):
    def __init__(self, cin, cout, up, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, 1e-6) for i in range(3)])
        self.upsamplers = nn.ModuleList([_Up(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.upsamplers[0](x) if self.upsamplers is not None else x


class _Up(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Down(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class VaeDown(nn.Module):
    def __init__(self, cin, cout, down, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, 1e-6) for i in range(2)])
        self.downsamplers = nn.ModuleList([_Down(cout)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.downsamplers[0](x) if self.downsamplers is not None else x


class Decoder(nn.Module):
    def __init__(self, c, latent, out_ch, groups):
        super().__init__()
        rev = c[::-1]
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = VaeMid(rev[0], groups)
        self.up_blocks = nn.ModuleList()
        prev = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(VaeUp(prev, co, i < len(c) - 1, groups))
            prev = co
        self.conv_norm_out = nn.GroupNorm(groups, c[0], eps=1e-6)
        self.conv_out = nn.Conv2d(c[0], out_ch, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for u in self.up_blocks:
            x = u(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Encoder(nn.Module):
    def __init__(self, c, latent, in_ch, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_ch, c[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        prev = c[0]
        for i, co in enumerate(c):
            self.down_blocks.append(VaeDown(prev, co, i < len(c) - 1, groups))
            prev = co
        self.mid_block = VaeMid(c[-1], groups)
        self.conv_norm_out = nn.GroupNorm(groups, c[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(c[-1], 2 * latent, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for d in self.down_blocks:
            x = d(x)
        return self.conv_out(F.silu(self.conv_norm_out(self.mid_block(x))))


class AutoencoderKLOracle(nn.Module):
    def __init__(self, block_out_channels=(128, 256, 512, 512), latent_channels=4, groups=32):
        super().__init__()
        c = list(block_out_channels)
        self.config = dict(block_out_channels=tuple(c), latent_channels=latent_channels)
        self.encoder = Encoder(c, latent_channels, 3, groups)
        self.decoder = Decoder(c, latent_channels, 3, groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))

    def encode_moments(self, x):
        """(mean, logvar) of DiagonalGaussianDistribution; sample = mean + exp(0.5 logvar) * eps"""
        mean, logvar = self.quant_conv(self.encoder(x)).chunk(2, dim=1)
        return mean, logvar.clamp(-30.0, 20.0)


def u8_from_image_fp16(image_fp16: torch.Tensor):
    """riffusion/riffusion_pipeline.py:430-434 on the reference's fp16 CUDA path, restated on the host:
    `(image / 2 + 0.5).clamp(0, 1)` on an fp16 tensor (one fp16 rounding per op), `.cpu().permute(0, 2, 3, 1).numpy()`
    (a float16 array), then DiffusionPipeline.numpy_to_pil `(images * 255).round().astype("uint8")` in float16.
    image_fp16: (B, 3, H, W) torch.float16 -> (B, H, W, 3) uint8 numpy."""
    assert image_fp16.dtype == torch.float16
    x = image_fp16.detach().cpu()
    x = ((x / 2) + 0.5).clamp(0, 1)
    arr = x.permute(0, 2, 3, 1).numpy()
    assert arr.dtype.name == "float16"
    return (arr * 255).round().astype("uint8")
