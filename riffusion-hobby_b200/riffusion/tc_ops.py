"""Thin tensor wrappers over the tensor-core C-ABI entry points (rf_gemm_f16, rf_conv2d_f16, ...).

Activations are fp16, NHWC for images and (rows, channels) for token matrices.  These helpers only
marshal pointers/strides; every FLOP runs in the tcgen05 kernels of librf_b200.so.
"""
from __future__ import annotations

import ctypes as C
import typing as T

import torch

from riffusion import _native

ACT_NONE, ACT_SILU = 0, 1


def _f16(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.float16:
        raise _native.NativeError(f"{name} must be a CUDA fp16 tensor (got {t.dtype} on {t.device})")
    return t


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def gemm(
    a: torch.Tensor, b: torch.Tensor, *, bias: T.Optional[torch.Tensor] = None, bias_per_row: bool = False,
    residual: T.Optional[torch.Tensor] = None, alpha: float = 1.0, act: int = ACT_NONE,
    out: T.Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.float16,
) -> torch.Tensor:
    """D[..., m, n] = act(alpha * A[..., m, :] . B[..., n, :] + bias) + residual.

    a: (..., M, K), b: (..., N, K) with up to two leading batch dims (strided views are fine as long
    as the last dim is contiguous and pitches are multiples of 8 elements).
    """
    _f16(a, "a"), _f16(b, "b")
    while a.dim() < 4:
        a = a.unsqueeze(0)
    while b.dim() < 4:
        b = b.unsqueeze(0)
    B2, B1, M, K = a.shape
    N = b.shape[2]
    b = b.expand(B2, B1, N, K)
    assert b.shape[3] == K and a.stride(3) == 1 and b.stride(3) == 1
    if out is None:
        out = torch.empty((B2, B1, M, N), dtype=out_dtype, device=a.device)
    o4 = out
    while o4.dim() < 4:
        o4 = o4.unsqueeze(0)
    assert o4.shape == (B2, B1, M, N) and o4.stride(3) == 1
    d = _native.GemmDesc()
    d.M, d.N, d.K, d.batch1, d.batch2 = M, N, K, B1, B2
    d.A, d.lda, d.sa1, d.sa2 = a.data_ptr(), a.stride(2), a.stride(1), a.stride(0)
    d.B, d.ldb, d.sb1, d.sb2 = b.data_ptr(), b.stride(2), b.stride(1), b.stride(0)
    d.D, d.ldd, d.sd1, d.sd2 = o4.data_ptr(), o4.stride(2), o4.stride(1), o4.stride(0)
    d.bias = None if bias is None else _f16(bias, "bias").data_ptr()
    d.bias_mode = 0 if bias is None else (2 if bias_per_row else 1)
    if residual is not None:
        r4 = _f16(residual, "residual")
        while r4.dim() < 4:
            r4 = r4.unsqueeze(0)
        assert r4.shape == (B2, B1, M, N) and r4.stride(3) == 1
        d.residual, d.ldr, d.sr1, d.sr2 = r4.data_ptr(), r4.stride(2), r4.stride(1), r4.stride(0)
    d.alpha, d.act, d.out_f32 = float(alpha), int(act), int(o4.dtype == torch.float32)
    with torch.cuda.device(a.device):
        _native.check(_native.lib().rf_gemm_f16(C.byref(d), _stream(a)))
    return out


def pack_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """torch Conv2d weight (Cout, Cin, kh, kw) -> (Cout, kh, kw, Cin) fp16 contiguous, the K-major
    layout the implicit-GEMM kernel streams with TMA."""
    return w.permute(0, 2, 3, 1).contiguous().to(torch.float16)


def conv2d(
    x: torch.Tensor, w_packed: torch.Tensor, *, x2: T.Optional[torch.Tensor] = None,
    bias: T.Optional[torch.Tensor] = None, bias_per_image: T.Optional[torch.Tensor] = None,
    residual: T.Optional[torch.Tensor] = None, stride: int = 1, act: int = ACT_NONE,
) -> torch.Tensor:
    """x (and optional x2, concatenated along channels): (B, H, W, C) fp16 NHWC contiguous.
    w_packed: (Cout, k, k, C1+C2).  Returns (B, Ho, Wo, Cout)."""
    _f16(x, "x"), _f16(w_packed, "w")
    B, H, W, C1 = x.shape
    C2 = 0 if x2 is None else x2.shape[3]
    Cout, k, _, Cin = w_packed.shape
    assert Cin == C1 + C2 and x.is_contiguous() and w_packed.is_contiguous()
    pad = 1 if k == 3 else 0
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = torch.empty((B, Ho, Wo, Cout), dtype=torch.float16, device=x.device)
    d = _native.ConvDesc()
    d.B, d.H, d.W, d.C1, d.C2, d.Cout, d.ksize, d.stride = B, H, W, C1, C2, Cout, k, stride
    d.x1 = x.data_ptr()
    d.x2 = None if x2 is None else _f16(x2, "x2").contiguous().data_ptr()
    d.w = w_packed.data_ptr()
    d.bias = None if bias is None else _f16(bias, "bias").data_ptr()
    d.bias_per_image = None if bias_per_image is None else _f16(bias_per_image, "bias_per_image").contiguous().data_ptr()
    if residual is not None:
        assert residual.shape == out.shape and residual.is_contiguous()
        d.residual = _f16(residual, "residual").data_ptr()
    d.out, d.alpha, d.act = out.data_ptr(), 1.0, int(act)
    with torch.cuda.device(x.device):
        _native.check(_native.lib().rf_conv2d_f16(C.byref(d), _stream(x)))
    return out
