"""Self-attention launches at the UNet's 64x64 level for ncu: python scratch/attn_one.py [B] [N] [d]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "riffusion-hobby_b200"))
import torch
from riffusion import tc_ops as ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
d = int(sys.argv[3]) if len(sys.argv) > 3 else 40
Nk = int(sys.argv[4]) if len(sys.argv) > 4 else N
heads = 8
C = heads * d
q = torch.randn(B, N, C, device="cuda").half()
k = torch.randn(B, Nk, C, device="cuda").half()
pitch = (Nk + 7) // 8 * 8
vt = torch.zeros(B, C, pitch, device="cuda").half()
vt[..., :Nk] = torch.randn(B, C, Nk, device="cuda").half()
for _ in range(3):
    o = ops.attention(q, k, vt, heads, Nk)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10):
    o = ops.attention(q, k, vt, heads, Nk)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
fl = 4.0 * B * heads * N * Nk * d
print(f"B={B} N={N} Nk={Nk} d={d}: {ms*1e3:.1f} us, {fl/ms/1e9:.0f} TFLOP/s (useful 4*N*N*d flops)")
