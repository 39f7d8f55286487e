"""One CFG UNet evaluation at bench.py's batch (B requests -> 2B images).
  python scratch/prof_eval.py 32            : ms per evaluation (CUDA graph replay, CUDA events) + eager kernel-class table
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X \
      python scratch/prof_eval.py 32 ncu    : launch list of exactly one eager evaluation (cudaProfilerStart/Stop)
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "riffusion-hobby_b200"))
import torch
from riffusion.graphed import GraphedUNet
from riffusion.riffusion_pipeline import RiffusionPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mode = sys.argv[2] if len(sys.argv) > 2 else "time"
pipe = RiffusionPipeline.random_init(seed=0, device="cuda", with_vae=False)
torch.manual_seed(0)
lat = torch.randn(B, 4, 64, 64, device="cuda").half()
ctx = torch.randn(2 * B, 77, 768, device="cuda").half()
x = torch.cat([lat, lat])
cache = {}
for _ in range(2):
    pipe.unet(x, 741, encoder_hidden_states=ctx, ctx_cache=cache)
torch.cuda.synchronize()
if mode == "ncu":
    torch.cuda.cudart().cudaProfilerStart()
    pipe.unet(x, 741, encoder_hidden_states=ctx, ctx_cache=cache)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
g = GraphedUNet(pipe.unet, lat.shape, ctx)
for _ in range(3):
    g(lat, 741)
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    g(lat, 741)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
tf = 2 * B * 0.803
print(f"CFG evaluation, {B} requests ({2*B} images): {ms:.3f} ms per evaluation (graph) = {tf / ms * 1e3:.0f} TFLOP/s algorithmic; "
      f"50 evals -> {50 * ms / B:.1f} ms/clip -> {B / (50 * ms) * 1e3:.2f} clips/s (UNet only)")
