import sys, os, time, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "riffusion-hobby_b200"))
import torch
from riffusion import _native
from riffusion.riffusion_pipeline import RiffusionPipeline
from riffusion.graphed import GraphedUNet
t0 = time.time()
pipe = RiffusionPipeline.random_init(seed=0, device="cuda", with_vae=True)
torch.cuda.synchronize(); print("init s", time.time() - t0)
lib = _native.lib()
def tc_prof(fn):
    lib.rf_tc_profile_begin()
    fn()
    ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_long()
    lib.rf_tc_profile_end(ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n))
    return ms.value, fl.value, n.value
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B in (1, 4, 8):
    x = torch.randn(2 * B, 4, 64, 64, device="cuda").half()
    ctx = torch.randn(2 * B, 77, 768, device="cuda").half()
    cache = {}
    f = lambda: pipe.unet(x, 741, encoder_hidden_states=ctx, ctx_cache=cache)
    eager = timeit(f, 3)
    ms, fl, n = tc_prof(f)
    g = GraphedUNet(pipe.unet, (B, 4, 64, 64), ctx)
    lat = x[:B]
    graph = timeit(lambda: g(lat, 741), 5)
    print(f"B={B}: eager {eager:.2f} ms, graph {graph:.2f} ms per CFG eval ({graph/B:.2f} ms/clip-eval); tc kernels {ms:.2f} ms over {n} launches, {fl/1e12:.3f} TFLOP -> {fl/ms/1e9:.1f} TFLOP/s in-kernel, {fl/graph/1e9:.1f} TFLOP/s per graph step")
    del g
z = torch.randn(4, 4, 64, 64, device="cuda").half()
f = lambda: pipe.vae.decode(z, scale=1 / 0.18215)
dec = timeit(f, 3)
ms, fl, n = tc_prof(f)
print(f"vae decode B=4: {dec:.2f} ms ({dec/4:.2f}/img); tc {ms:.2f} ms, {fl/1e12:.3f} TFLOP -> {fl/ms/1e9:.1f} TFLOP/s")
