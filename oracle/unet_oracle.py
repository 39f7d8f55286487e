"""Oracle for hot path (b): a plain-PyTorch restatement of the Stable-Diffusion-1.5 UNet2DConditionModel,
the PNDM (PLMS) scheduler and the img2img denoising loop of `RiffusionPipeline.interpolate_img2img`.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED.  The arithmetic lives in `diffusers==0.9.0` (requirements.txt:5), which is not installed,
not in /opt/wheelhouse and not under /root/reference; no reference test touches this path.  The
architecture below is restated from the published SD-1.5 configuration (unet/config.json of
runwayml/stable-diffusion-v1-5: block_out_channels (320,640,1280,1280), layers_per_block 2,
attention_head_dim 8, cross_attention_dim 768, norm_num_groups 32, flip_sin_to_cos, freq_shift 0) and from
memory of diffusers 0.9's module structure; parameter names follow the diffusers state_dict so a real
checkpoint loads with `load_state_dict`.  What IS anchored on the reference is the control flow around it:
riffusion/riffusion_pipeline.py:289-436 (call sites :314, :361-365, :379, :392-396, :403, :406-408,
:411-415, :418, :421-425) and riffusion/util/torch_util.py:21-48 (slerp).
Self-checks that pin the restatement as far as possible: parameter count == 859,520,964 (the published
860 M), alphas_cumprod[0] / [999] == 0.99915 / 0.0046601 (the well-known SD constants), PLMS timestep
table and start indices of SURVEY Appendix B.
"""
from __future__ import annotations

import math
import typing as T

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------ UNet
class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb: T.Optional[int] = 1280, groups: int = 32, eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout) if temb else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class CrossAttention(nn.Module):
    def __init__(self, query_dim: int, context_dim: T.Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        context = x if context is None else context
        B, Nq, _ = x.shape
        h = self.heads
        q = self.to_q(x).view(B, Nq, h, -1).transpose(1, 2)
        k = self.to_k(context).view(B, context.shape[1], h, -1).transpose(1, 2)
        v = self.to_v(context).view(B, context.shape[1], h, -1).transpose(1, 2)
        attn = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * self.scale, dim=-1)
        out = torch.matmul(attn, v).transpose(1, 2).reshape(B, Nq, -1)
        return self.to_out[0](out)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, context_dim: int):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, heads, dim_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, x, context):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    def __init__(self, channels: int, heads: int, context_dim: int, groups: int = 32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.proj_in = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, channels // heads, context_dim)])
        self.proj_out = nn.Conv2d(channels, channels, 1)

    def forward(self, x, context):
        B, C, H, W = x.shape
        h = self.proj_in(self.norm(x)).permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            h = blk(h, context)
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
        return self.proj_out(h) + x


class Downsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, heads, ctx, attn: bool, down: bool, groups: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups) for i in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, ctx, groups) for _ in range(2)]) if attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, c, temb, heads, ctx, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, groups), ResnetBlock2D(c, c, temb, groups)])
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, ctx, groups)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, prev, cout, skips: T.Sequence[int], temb, heads, ctx, attn: bool, up: bool, groups: int):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D((prev if i == 0 else cout) + skips[i], cout, temb, groups) for i in range(3)])
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, ctx, groups) for _ in range(3)]) if attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if up else None

    def forward(self, x, skips, temb, ctx):
        for i, r in enumerate(self.resnets):
            x = r(torch.cat([x, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1, self.linear_2 = nn.Linear(cin, cout), nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


def timestep_sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0, scale=1, max_period=10000)"""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class UNet2DConditionOracle(nn.Module):
    """SD-1.5 UNet.  `block_out_channels` etc. can be shrunk for fast unit tests."""

    def __init__(self, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), heads=8,
                 cross_attention_dim=768, groups=32):
        super().__init__()
        c = list(block_out_channels)
        temb = c[0] * 4
        self.config = dict(in_channels=in_channels, out_channels=out_channels, block_out_channels=tuple(c), heads=heads,
                           cross_attention_dim=cross_attention_dim, groups=groups)
        self.conv_in = nn.Conv2d(in_channels, c[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(c[0], temb)
        self.down_blocks = nn.ModuleList()
        skip_ch = [c[0]]
        prev = c[0]
        for i, co in enumerate(c):
            last = i == len(c) - 1
            self.down_blocks.append(DownBlock(prev, co, temb, heads, cross_attention_dim, attn=not last, down=not last, groups=groups))
            skip_ch += [co, co] + ([] if last else [co])
            prev = co
        self.mid_block = MidBlock(c[-1], temb, heads, cross_attention_dim, groups)
        self.up_blocks = nn.ModuleList()
        rev = c[::-1]
        for i, co in enumerate(rev):
            skips = [skip_ch.pop() for _ in range(3)]
            self.up_blocks.append(UpBlock(prev, co, skips, temb, heads, cross_attention_dim, attn=i > 0,
                                          up=i < len(c) - 1, groups=groups))
            prev = co
        self.conv_norm_out = nn.GroupNorm(groups, c[0], eps=1e-5)
        self.conv_out = nn.Conv2d(c[0], out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states):
        t = torch.as_tensor(timestep, device=sample.device).reshape(-1).expand(sample.shape[0])
        temb = self.time_embedding(timestep_sinusoid(t, self.config["block_out_channels"][0]).to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states)
            skips += outs
        x = self.mid_block(x, temb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


def init_weights_(model: nn.Module, seed: int = 0, std: float = 0.02) -> nn.Module:
    """BASELINE config 4: random-init SD-1.5 weights N(0, 0.02^2), norm scales 1, biases 0 — plus a small
    non-zero bias / non-unit norm scale so that bias and affine paths are actually exercised by parity tests."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g) * std)
            elif name.endswith("weight"):       # norm scales
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:                               # biases
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    return model


# ------------------------------------------------------------------------------------------ scheduler
class PNDMSchedulerOracle:
    """diffusers PNDMScheduler(skip_prk_steps=True, steps_offset=1, scaled_linear betas 0.00085..0.012,
    set_alpha_to_one=False) — SURVEY Appendix B [memory]."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.config = {"steps_offset": steps_offset}
        self.timesteps: T.Optional[torch.Tensor] = None

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        step_ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * step_ratio).round() + self.config["steps_offset"]
        plms = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets: T.List[torch.Tensor] = []
        self.counter = 0
        self.cur_sample = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original, noise, timestep):
        a = self.alphas_cumprod[int(timestep)]
        return a.sqrt().to(original.dtype) * original + (1 - a).sqrt().to(original.dtype) * noise

    def coefficients(self, timestep: int, prev_timestep: int) -> T.Tuple[float, float]:
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t, b_p = 1 - a_t, 1 - a_p
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return float(sample_coeff), float((a_p - a_t) / denom)

    def step(self, model_output, timestep: int, sample):
        timestep = int(timestep)
        prev = timestep - self.num_train_timesteps // self.num_inference_steps
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev = timestep
            timestep = timestep + self.num_train_timesteps // self.num_inference_steps
        if len(self.ets) == 1 and self.counter == 0:
            e = model_output
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            e = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            e = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            e = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            e = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        ca, cb = self.coefficients(timestep, prev)
        self.counter += 1
        return ca * sample - cb * e


def slerp(t: float, v0: torch.Tensor, v1: torch.Tensor, dot_threshold: float = 0.9995) -> torch.Tensor:
    """riffusion/util/torch_util.py:21-48 — numpy in the tensors' dtype."""
    a, b = v0.detach().cpu().numpy(), v1.detach().cpu().numpy()
    dot = np.sum(a * b / (np.linalg.norm(a) * np.linalg.norm(b)))
    if np.abs(dot) > dot_threshold:
        out = (1 - t) * a + t * b
    else:
        th0 = np.arccos(dot)
        s0 = np.sin(th0 - th0 * t) / np.sin(th0)
        s1 = np.sin(th0 * t) / np.sin(th0)
        out = s0 * a + s1 * b
    return torch.from_numpy(np.asarray(out)).to(v0.device)


def img2img_loop(unet, scheduler: PNDMSchedulerOracle, text_embeddings, uncond_embeddings, init_latents, noise_a,
                 noise_b, alpha: float, strength: float, num_inference_steps: int, guidance_scale: float,
                 mask=None) -> T.Tuple[torch.Tensor, int]:
    """riffusion/riffusion_pipeline.py:311-425 with the tensors the reference draws from its generators
    (noise_a/noise_b, :371-376) injected.  Returns (latents, number of UNet evaluations)."""
    scheduler.set_timesteps(num_inference_steps)
    ctx = torch.cat([uncond_embeddings, text_embeddings])                       # :354
    offset = scheduler.config.get("steps_offset", 0)                            # :361
    init_timestep = min(int(num_inference_steps * strength) + offset, num_inference_steps)   # :362-363
    t0 = int(scheduler.timesteps[-init_timestep])                               # :365
    noise = slerp(float(alpha), noise_a, noise_b)                               # :377
    init_orig = init_latents
    latents = scheduler.add_noise(init_latents, noise, t0)                      # :379
    t_start = max(num_inference_steps - init_timestep + offset, 0)              # :392
    n_evals = 0
    for t in scheduler.timesteps[t_start:]:                                     # :396-398
        x2 = scheduler.scale_model_input(torch.cat([latents] * 2), t)           # :400-403
        eps = unet(x2, int(t), ctx)                                             # :406-408
        n_evals += 1
        eu, et = eps.chunk(2)
        eps = eu + guidance_scale * (et - eu)                                   # :411-415
        latents = scheduler.step(eps, int(t), latents)                          # :418
        if mask is not None:                                                    # :420-425
            proper = scheduler.add_noise(init_orig, noise, int(t))
            latents = proper * mask + latents * (1 - mask)
    return latents, n_evals
