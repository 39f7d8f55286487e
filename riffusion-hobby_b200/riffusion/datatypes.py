"""Request / response schema of the inference API (same fields and defaults as the reference's
riffusion/datatypes.py:10-73, which the Flask server fills with dacite)."""
from __future__ import annotations

import typing as T
from dataclasses import dataclass


@dataclass(frozen=True)
class PromptInput:
    """One end point of an interpolation: text prompt, RNG seed and per-prompt sampler settings."""

    prompt: str
    seed: int
    negative_prompt: T.Optional[str] = None   # carried in the schema; `riffuse` never forwards it (reference quirk)
    denoising: float = 0.75                    # img2img strength
    guidance: float = 7.0                      # classifier-free guidance scale


@dataclass(frozen=True)
class InferenceInput:
    """A (start, end, alpha) interpolation request on a seed spectrogram image."""

    start: PromptInput
    end: PromptInput
    alpha: float                               # 0 = start, 1 = end
    num_inference_steps: int = 50
    seed_image_id: str = "og_beat"
    mask_image_id: T.Optional[str] = None


@dataclass(frozen=True)
class InferenceOutput:
    """Response of the model server: base64 JPEG image, base64 MP3 audio, clip duration."""

    image: str
    audio: str
    duration_s: float
