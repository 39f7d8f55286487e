"""Device checks and slerp (reference: riffusion/util/torch_util.py)."""
from __future__ import annotations

import warnings

import numpy as np
import torch


def check_device(device: str, backup: str = "cpu") -> str:
    """Validate a device string.

    The reference silently falls back to `backup` with a warning when CUDA/MPS is missing
    (util/torch_util.py:7-18).  The B200 build has no CPU path for the hot kernels, so a
    missing CUDA device is an error here; the `mps` warning text is kept for API parity.
    """
    dev = device.lower()
    if dev.startswith("cuda") and not torch.cuda.is_available():
        raise RuntimeError(
            f"{device} is not available and the B200-native riffusion build has no CPU fallback "
            f"(the reference would have warned and used {backup})"
        )
    if dev.startswith("mps"):
        warnings.warn(f"WARNING: {device} is not available, using {backup} instead.", stacklevel=3)
        return backup
    return device


def slerp(t: float, v0: torch.Tensor, v1: torch.Tensor, dot_threshold: float = 0.9995) -> torch.Tensor:
    """Spherical interpolation between two tensors, bit-compatible host-numpy mode.

    Restates util/torch_util.py:21-48: everything happens in numpy *in the tensors' dtype*
    (fp16 on GPU runs), falling back to lerp when |cos| > dot_threshold.  The fused on-device
    version used by the denoising loop lives in the native library; this one is the parity
    reference for it and the API-compatible entry point.
    """
    if not isinstance(v0, torch.Tensor):
        raise TypeError("slerp expects torch tensors (the reference's numpy branch is broken: "
                        "`inputs_are_torch` is unset, util/torch_util.py:27-45)")
    device = v0.device
    a = v0.detach().cpu().numpy()
    b = v1.detach().cpu().numpy()
    dot = np.sum(a * b / (np.linalg.norm(a) * np.linalg.norm(b)))
    if np.abs(dot) > dot_threshold:
        out = (1 - t) * a + t * b
    else:
        theta_0 = np.arccos(dot)
        sin_theta_0 = np.sin(theta_0)
        theta_t = theta_0 * t
        s0 = np.sin(theta_0 - theta_t) / sin_theta_0
        s1 = np.sin(theta_t) / sin_theta_0
        out = s0 * a + s1 * b
    return torch.from_numpy(np.asarray(out)).to(device)
