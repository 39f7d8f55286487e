"""One Griffin-Lim step for ncu captures: python scratch/prof_gl.py [B] [n_iter]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "riffusion-hobby_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from riffusion.spectrogram_converter import SpectrogramConverter
from riffusion.spectrogram_params import SpectrogramParams
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
conv = SpectrogramConverter(SpectrogramParams(num_griffin_lim_iters=n_iter), "cuda")
mel = bench.synthetic_mel(B, 0).cuda()
ang = torch.rand((B, 8821, 512), dtype=torch.complex64, device="cuda")
w = conv.waveform_from_mel_amplitudes(mel, ang)
torch.cuda.synchronize()
print("ok", w.shape, float(w.abs().max()))
