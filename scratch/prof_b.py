import sys, os, ctypes, collections, csv
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "riffusion-hobby_b200"))
import torch
from riffusion import _native
from riffusion.riffusion_pipeline import RiffusionPipeline
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pipe = RiffusionPipeline.random_init(seed=0, device="cuda", with_vae=False)
lib = _native.lib()
x = torch.randn(2 * B, 4, 64, 64, device="cuda").half(); ctx = torch.randn(2 * B, 77, 768, device="cuda").half()
cache = {}
for _ in range(2): pipe.unet(x, 741, encoder_hidden_states=ctx, ctx_cache=cache)
torch.cuda.synchronize()
os.environ["RF_TC_PROFILE_DUMP"] = "gpurun_out/tc_launches.csv"
os.makedirs("gpurun_out", exist_ok=True)
lib.rf_tc_profile_begin()
pipe.unet(x, 741, encoder_hidden_states=ctx, ctx_cache=cache)
ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_long()
lib.rf_tc_profile_end(ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n))
rows = list(csv.DictReader(open("gpurun_out/tc_launches.csv")))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    k = (r["conv"], r["M"], r["N"], r["K"], r["batch"], r.get("splits", "1"), r.get("bn", "0"))
    a = agg[k]; a[0] += 1; a[1] += float(r["ms"]); a[2] += 2.0 * int(r["M"]) * int(r["N"]) * int(r["K"]) * int(r["batch"])
print(f"total {ms.value:.2f} ms, {fl.value/1e12:.3f} TFLOP, {n.value} launches")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"conv={k[0]} M={k[1]:>6} N={k[2]:>5} K={k[3]:>6} batch={k[4]:>3} S={k[5]} bn={k[6]:>4}  n={a[0]:3d}  {a[1]:7.3f} ms  {a[2]/a[1]/1e9:8.1f} TFLOP/s")
