"""Oracle for hot path (a): a CPU restatement of the reference's audio arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference (riffusion/spectrogram_converter.py) delegates all arithmetic to torchaudio
transforms; torchaudio in turn calls torch.stft / torch.istft / torch.linalg.lstsq (ATen, no
source on this box).  This module restates that arithmetic explicitly in numpy/torch-CPU —
framing, windowing, rFFT, overlap-add, envelope division, the Griffin-Lim recurrence, the
triangular mel filterbank and the minimum-norm inverse — with each function citing the file:line
it follows.  `TA/` = site-packages/torchaudio (2.11.0).

Pinning (tests/test_oracle_cpu.py): every function here is checked against the installed
torchaudio/torch ops themselves (the reference's actual third-party arithmetic, importable on
CPU), and the forward path is checked against the reference's own fixture pair
test/test_data/tired_traveler/{clips/clip_2*.wav, images/clip_2*_stereo.png} including the EXIF
MAX_VALUE (committed as tests/golden/tired_traveler_clip2.npz by tests/golden/make_golden.py).
The reference's tests hold no numeric vector for Griffin-Lim / inverse mel; those are pinned
against torchaudio with injected initial phases.
"""
from __future__ import annotations

import math
import typing as T

import numpy as np
import torch


# --------------------------------------------------------------------------- parameters
def derived_sizes(sample_rate=44100, step_size_ms=10, window_duration_ms=100, padded_duration_ms=400):
    """n_fft, win_length, hop_length — riffusion/spectrogram_params.py:62-81."""
    n_fft = int(padded_duration_ms / 1000.0 * sample_rate)
    win = int(window_duration_ms / 1000.0 * sample_rate)
    hop = int(step_size_ms / 1000.0 * sample_rate)
    return n_fft, win, hop


def hann_window(win_length: int, dtype=torch.float32) -> torch.Tensor:
    """torch.hann_window (periodic) as passed at riffusion/spectrogram_converter.py:52,66.
    The fp32 window torch builds differs from the exactly-rounded 0.5-0.5cos(2 pi n/win) by up to
    ~3e-8 absolute (fp32 cosine), so the oracle takes torch's values: that is the window the
    reference multiplies by."""
    return torch.hann_window(win_length, periodic=True, dtype=dtype)


# --------------------------------------------------------------------------- mel filterbank
def melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, norm=None, mel_scale="htk") -> torch.Tensor:
    """TA/functional/functional.py:518-587 (htk: :440, :474; triangles :507-513), fp32 ops in the
    same order so the result is bit-identical to torchaudio's buffer."""
    assert mel_scale == "htk", "oracle restates the htk scale the reference defaults to"
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down_slopes, up_slopes))
    if norm == "slaney":
        fb *= (2.0 / (f_pts[2 : n_mels + 2] - f_pts[:n_mels])).unsqueeze(0)
    return fb


# --------------------------------------------------------------------------- STFT / iSTFT
def stft(x: np.ndarray, n_fft: int, hop: int, win: np.ndarray, dtype=np.float64) -> np.ndarray:
    """torch.stft(center=True, pad_mode="reflect", onesided, not normalized) as called from
    TA/functional/functional.py:123-134 (Spectrogram, riffusion/spectrogram_converter.py:47-59).
    x: (B, L) -> (B, n_fft//2+1, 1 + L//hop) complex.  The window (win_length <= n_fft) is
    zero-padded centred to n_fft."""
    x = np.asarray(x, dtype=dtype)
    B, L = x.shape
    if L <= n_fft // 2:
        raise ValueError("Padding size should be less than the corresponding input dimension")
    W = len(win)
    left = (n_fft - W) // 2
    wpad = np.zeros(n_fft, dtype=dtype)
    wpad[left : left + W] = win
    xp = np.pad(x, ((0, 0), (n_fft // 2, n_fft // 2)), mode="reflect")
    T_ = 1 + L // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T_)[:, None]
    frames = xp[:, idx] * wpad  # (B, T, n_fft)
    spec = np.fft.rfft(frames, axis=-1)
    return np.ascontiguousarray(np.transpose(spec, (0, 2, 1)))


def istft(spec: np.ndarray, n_fft: int, hop: int, win: np.ndarray, dtype=np.float64) -> np.ndarray:
    """torch.istft(center=True, length=None) as called at TA/functional/functional.py:318-320:
    irfft per frame, multiply by the padded window, overlap-add, divide by the overlap-added
    squared window, trim n_fft//2 on both sides.  spec: (B, F, T) -> (B, hop*(T-1))."""
    B, F, T_ = spec.shape
    W = len(win)
    left = (n_fft - W) // 2
    wpad = np.zeros(n_fft, dtype=dtype)
    wpad[left : left + W] = win
    frames = np.fft.irfft(np.transpose(spec, (0, 2, 1)), n=n_fft, axis=-1) * wpad  # (B, T, n_fft)
    total = n_fft + hop * (T_ - 1)
    y = np.zeros((B, total), dtype=dtype)
    env = np.zeros(total, dtype=dtype)
    for t in range(T_):
        y[:, t * hop : t * hop + n_fft] += frames[:, t]
        env[t * hop : t * hop + n_fft] += wpad * wpad
    start, end = n_fft // 2, total - n_fft // 2
    return y[:, start:end] / env[start:end]


def griffinlim(
    specgram: np.ndarray, n_fft: int, hop: int, win: np.ndarray, n_iter: int, momentum: float,
    init_angles: T.Optional[np.ndarray], power: float = 1.0, dtype=np.float64,
) -> np.ndarray:
    """F.griffinlim, TA/functional/functional.py:255-353 with length=None.
    init_angles: (B, F, T) complex = the `torch.rand(cfloat)` draw of :310, or None for
    rand_init=False (:312).  Returns (B, hop*(T-1))."""
    if not 0 <= momentum < 1:
        raise ValueError("momentum must be in range [0, 1). Found: {}".format(momentum))
    cdtype = np.complex128 if dtype == np.float64 else np.complex64
    momentum = momentum / (1 + momentum)                                   # :300
    spec = np.asarray(specgram, dtype=dtype) ** (1 / power)                # :306
    angles = np.ones(spec.shape, cdtype) if init_angles is None else np.asarray(init_angles, cdtype)
    tprev = 0.0                                                            # :315
    for _ in range(n_iter):
        inverse = istft(spec * angles, n_fft, hop, win, dtype)            # :318-320
        rebuilt = stft(inverse, n_fft, hop, win, dtype).astype(cdtype)    # :323-334
        angles = rebuilt
        if momentum:
            angles = angles - tprev * dtype(momentum)                      # :337-339
        angles = angles / (np.abs(angles) + dtype(1e-16))                  # :340
        tprev = rebuilt                                                    # :343
    return istft(spec * angles, n_fft, hop, win, dtype)                    # :346-348


# --------------------------------------------------------------------------- mel / inverse mel
def mel_scale(spec_mag: np.ndarray, fb: np.ndarray) -> np.ndarray:
    """MelScale.forward, TA/transforms/_transforms.py:417: (.., F, T) x (F, M) -> (.., M, T)."""
    return np.swapaxes(np.swapaxes(spec_mag, -1, -2) @ fb, -1, -2)


def inverse_mel(mel: np.ndarray, fb: np.ndarray) -> np.ndarray:
    """InverseMelScale.forward, TA/transforms/_transforms.py:508:
    relu(lstsq(fb^T (M x F), mel (M x T), driver="gels").solution).  For the wide full-row-rank
    fb^T LAPACK gels returns the minimum-norm solution fb (fb^T fb)^-1 mel; computed in fp64."""
    fb64 = np.asarray(fb, np.float64)
    gram = fb64.T @ fb64
    sol = fb64 @ np.linalg.solve(gram, np.asarray(mel, np.float64))
    return np.maximum(sol, 0.0)


def mel_amplitudes_from_waveform(wave: np.ndarray, fb: np.ndarray, n_fft, hop, win, dtype=np.float64):
    """SpectrogramConverter.mel_amplitudes_from_waveform, riffusion/spectrogram_converter.py:165-185."""
    return mel_scale(np.abs(stft(wave, n_fft, hop, win, dtype)), np.asarray(fb, dtype))


def waveform_from_mel_amplitudes(mel, fb, n_fft, hop, win, n_iter, init_angles, momentum=0.99, dtype=np.float64):
    """SpectrogramConverter.waveform_from_mel_amplitudes, riffusion/spectrogram_converter.py:187-204."""
    lin = inverse_mel(mel, fb).astype(dtype)
    return griffinlim(lin, n_fft, hop, win, n_iter, momentum, init_angles, dtype=dtype)


# --------------------------------------------------------------------------- image / pcm quantisation
def spectrogram_from_image_array(rgb: np.ndarray, power=0.25, stereo=False, max_value=30e6) -> np.ndarray:
    """image_util.spectrogram_from_image after PIL's P/L->RGB conversion,
    riffusion/util/image_util.py:84-110.  rgb: (H, W, 3) uint8 -> (C, H, W) float32."""
    data = rgb[::-1].transpose(2, 0, 1)                       # flip Y (:85), channels first (:88)
    data = data[[1, 2], :, :] if stereo else data[0:1, :, :]  # (:89-93)
    data = data.astype(np.float32)
    data = 255 - data
    data = data / 255
    data = np.power(data, 1 / power)
    data = data * max_value
    return data


def image_array_from_spectrogram(spectrogram: np.ndarray, power=0.25) -> T.Tuple[np.ndarray, float]:
    """image_util.image_from_spectrogram, riffusion/util/image_util.py:27-56.
    (C, H, W) float32 -> ((H, W, 3) uint8, max_value)."""
    max_value = np.max(spectrogram)
    data = spectrogram / max_value
    data = np.power(data, power)
    data = data * 255
    data = 255 - data
    data = data.astype(np.uint8)
    if data.shape[0] == 1:
        img = np.repeat(data[0][:, :, None], 3, axis=2)        # L -> RGB replicates the plane
    elif data.shape[0] == 2:
        img = np.array([np.zeros_like(data[0]), data[0], data[1]]).transpose(1, 2, 0)
    else:
        raise NotImplementedError(f"Unsupported number of channels: {data.shape[0]}")
    return img[::-1].copy(), float(max_value)


def int16_from_waveform(samples: np.ndarray, normalize=True) -> np.ndarray:
    """audio_util.audio_from_waveform up to the int16 cast, riffusion/util/audio_util.py:22-28."""
    samples = np.array(samples, dtype=np.float32, copy=True)
    if normalize:
        samples *= np.iinfo(np.int16).max / np.max(np.abs(samples))
    return samples.transpose(1, 0).astype(np.int16)
