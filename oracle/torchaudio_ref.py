"""The reference's own CPU implementation of hot path (a): installed torchaudio transforms,
constructed with exactly the arguments riffusion/spectrogram_converter.py:47-99 passes, minus the
torchaudio-0.13-only InverseMelScale kwargs (max_iter / tolerance_* / sgdargs, :93-96) that raise
TypeError on torchaudio 2.x — i.e. the `lstsq(gels)` InverseMelScale.

TEST / BENCH INFRASTRUCTURE ONLY: used as the parity checker in tests/ and as the timed CPU
baseline (`bench.py --impl reference`, `cpu_baseline.kind = "reference"`).
"""
from __future__ import annotations

import torch
import torchaudio


class TorchaudioConverter:
    """Same four transforms and the same two torch-only methods as the reference class."""

    def __init__(self, sample_rate=44100, n_fft=17640, win_length=4410, hop_length=441, n_mels=512,
                 f_min=0, f_max=10000, n_iter=32, mel_norm=None, mel_scale="htk", device="cpu"):
        self.device = device
        self.spectrogram_func = torchaudio.transforms.Spectrogram(   # spectrogram_converter.py:47-59
            n_fft=n_fft, hop_length=hop_length, win_length=win_length, pad=0,
            window_fn=torch.hann_window, power=None, normalized=False, wkwargs=None,
            center=True, pad_mode="reflect", onesided=True,
        ).to(device)
        self.inverse_spectrogram_func = torchaudio.transforms.GriffinLim(   # :61-73
            n_fft=n_fft, n_iter=n_iter, win_length=win_length, hop_length=hop_length,
            window_fn=torch.hann_window, power=1.0, wkwargs=None, momentum=0.99, length=None,
            rand_init=True,
        ).to(device)
        self.mel_scaler = torchaudio.transforms.MelScale(   # :75-84
            n_mels=n_mels, sample_rate=sample_rate, f_min=f_min, f_max=f_max,
            n_stft=n_fft // 2 + 1, norm=mel_norm, mel_scale=mel_scale,
        ).to(device)
        self.inverse_mel_scaler = torchaudio.transforms.InverseMelScale(   # :86-99 (2.x signature)
            n_stft=n_fft // 2 + 1, n_mels=n_mels, sample_rate=sample_rate, f_min=f_min, f_max=f_max,
            norm=mel_norm, mel_scale=mel_scale,
        ).to(device)

    def mel_amplitudes_from_waveform(self, waveform: torch.Tensor) -> torch.Tensor:   # :165-185
        return self.mel_scaler(torch.abs(self.spectrogram_func(waveform)))

    def waveform_from_mel_amplitudes(self, amplitudes_mel: torch.Tensor, init_angles=None) -> torch.Tensor:  # :187-204
        lin = self.inverse_mel_scaler(amplitudes_mel)
        if init_angles is None:
            return self.inverse_spectrogram_func(lin)
        return griffinlim_with_angles(self.inverse_spectrogram_func, lin, init_angles)


def griffinlim_with_angles(gl: torchaudio.transforms.GriffinLim, lin: torch.Tensor, init_angles: torch.Tensor):
    """Run torchaudio's own F.griffinlim but make its `torch.rand(..., dtype=cfloat)` draw
    (TA/functional/functional.py:310) return `init_angles`, so two implementations can be compared
    on identical initial phases."""
    orig = torch.rand
    packed = init_angles.reshape([-1] + list(init_angles.shape[-2:])).clone()

    def fake_rand(*a, **k):
        if k.get("dtype") in (torch.complex64, torch.complex128):
            return packed.to(k["dtype"])
        return orig(*a, **k)

    torch.rand = fake_rand
    try:
        return gl(lin)
    finally:
        torch.rand = orig
