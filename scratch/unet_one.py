"""Three eager CFG UNet evaluations at batch B (for an ncu launch list): python scratch/unet_one.py [B]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "riffusion-hobby_b200"))
import torch
from riffusion.riffusion_pipeline import RiffusionPipeline
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pipe = RiffusionPipeline.random_init(seed=0, device="cuda", with_vae=False)
x = torch.randn(2 * B, 4, 64, 64, device="cuda").half(); ctx = torch.randn(2 * B, 77, 768, device="cuda").half()
cache = {}
for _ in range(3):
    pipe.unet(x, 741, encoder_hidden_states=ctx, ctx_cache=cache)
torch.cuda.synchronize()
print("ok")
