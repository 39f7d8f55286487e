"""A few launches of one GEMM shape for ncu: python scratch/gemm_one.py M N K [geglu|res|plain]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "riffusion-hobby_b200"))
import torch
from riffusion import tc_ops as ops
M, N, K = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "plain"
a = torch.randn(M, K, device="cuda").half()
b = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
bias = torch.randn(N, device="cuda").half()
res = torch.randn(M, N, device="cuda").half() if mode == "res" else None
for _ in range(4):
    if mode == "geglu":
        out = ops.gemm(a, b, bias=bias, act=ops.ACT_GEGLU)
    else:
        out = ops.gemm(a, b, bias=bias, residual=res)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(20):
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GEGLU) if mode == "geglu" else ops.gemm(a, b, bias=bias, residual=res)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 20
print(f"M={M} N={N} K={K} {mode}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s")
