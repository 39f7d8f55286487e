import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "riffusion-hobby_b200"))
import torch, torchaudio, numpy as np
from riffusion.spectrogram_params import SpectrogramParams
from riffusion import spectrogram_converter as sc
p = SpectrogramParams()
dev = "cuda"
conv = sc.SpectrogramConverter(p, dev)
N, W, H, F = p.n_fft, p.win_length, p.hop_length, p.n_fft // 2 + 1
win = torch.hann_window(W)
torch.manual_seed(0)
# 1. stft
x = torch.randn(2, 30000) * 1000
ref = torch.stft(x, N, H, W, win, center=True, pad_mode="reflect", return_complex=True)
out = conv.spectrogram_func(x.cuda()).cpu()
print("stft rel err", ((out - ref).abs().max() / ref.abs().max()).item(), out.shape)
# 2. mel scale + fused stft_mel
fb = sc.mel_filterbank(F, 0.0, 10000.0, 512, 44100)
fb_ta = torchaudio.functional.melscale_fbanks(F, 0.0, 10000.0, 512, 44100, None, "htk")
print("fb bit-identical:", torch.equal(fb, fb_ta))
mel_ref = torch.matmul(ref.abs().transpose(-1, -2), fb_ta).transpose(-1, -2)
mel = conv.mel_amplitudes_from_waveform(x.cuda()).cpu()
print("stft_mel rel err", ((mel - mel_ref).abs().max() / mel_ref.abs().max()).item())
mel2 = conv.mel_scaler(ref.abs().cuda()).cpu()
print("mel_scale rel err", ((mel2 - mel_ref).abs().max() / mel_ref.abs().max()).item())
# 3. inverse mel
T = 64
melin = torch.rand(2, 512, T) * 1e6
inv_ref = torch.relu(torch.linalg.lstsq(fb_ta.transpose(-1, -2)[None], melin, driver="gels").solution)
inv = conv.inverse_mel_scaler(melin.cuda()).cpu()
print("inverse mel rel err", ((inv - inv_ref).norm() / inv_ref.norm()).item(), "max", ((inv - inv_ref).abs().max() / inv_ref.abs().max()).item())
# 4. griffin-lim vs torchaudio with injected angles
for T, n_iter in ((24, 2), (64, 32)):
    lin = inv_ref[:, :, :T].contiguous() if T <= 64 else None
    ang = torch.rand(2, F, T, dtype=torch.cfloat)
    orig = torch.rand
    torch.rand = lambda *a, **k: ang.clone()
    try:
        refw = torchaudio.functional.griffinlim(lin, win, N, H, W, 1.0, n_iter, 0.99, None, True)
    finally:
        torch.rand = orig
    p2 = SpectrogramParams(num_griffin_lim_iters=n_iter)
    c2 = sc.SpectrogramConverter(p2, dev)
    w = c2.inverse_spectrogram_func.forward(lin.cuda(), ang.cuda()).cpu()
    w2 = c2.waveform_from_mel_amplitudes(melin[:, :, :T].contiguous().cuda(), ang.cuda()).cpu()
    pk = refw.abs().max()
    print(f"GL T={T} it={n_iter}: rel rms {((w - refw).norm() / refw.norm()).item():.3e}  norm-rms {(((w - refw) / pk).pow(2).mean().sqrt()).item():.3e}"
          f" fused rel {((w2 - refw).norm() / refw.norm()).item():.3e}")
# 5. timing GL batch 64, T=512
B, T = 64, 512
mel = torch.rand(B, 512, T, device=dev) * 1e6
ang = torch.rand(B, F, T, dtype=torch.cfloat, device=dev)
for _ in range(2):
    wv = conv.waveform_from_mel_amplitudes(mel, ang)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(3):
    wv = conv.waveform_from_mel_amplitudes(mel, ang)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
gb = B * (32 * 36 * 4000 * 512 + 12 * 4000 * 512 + 65 * 4 * 441 * 511) / 1e9
print(f"GL B=64 T=512 32it: {ms:.2f} ms  -> {B / ms * 1e3:.1f} clips/s, algorithmic {gb:.1f} GB -> {gb / ms * 1e3:.0f} GB/s ({gb / ms * 1e3 / 6570.9:.3f} of measured HBM)")
print("finite:", torch.isfinite(wv).all().item())
