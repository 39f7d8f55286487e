"""Pin the oracle (oracle/audio_oracle.py) against the reference's third-party arithmetic
(installed torchaudio / torch) and against the reference's own fixture pair.  CPU only."""
import numpy as np
import pytest
import torch
import torchaudio

from oracle import audio_oracle as ao
from oracle.torchaudio_ref import TorchaudioConverter, griffinlim_with_angles

N, W, H = ao.derived_sizes()
F = N // 2 + 1


def test_derived_sizes():
    assert (N, W, H) == (17640, 4410, 441)
    assert ao.derived_sizes(sample_rate=48000) == (19200, 4800, 480)


def test_window_and_fbanks_bit_identical_to_torchaudio():
    w32 = ao.hann_window(W)
    n = torch.arange(W, dtype=torch.float64)
    exact = 0.5 - 0.5 * torch.cos(2 * torch.pi * n / W)
    assert torch.equal(w32, torch.hann_window(W))
    assert (w32.double() - exact).abs().max() < 3e-7      # fp32 cosine error of torch's window
    for (f0, f1) in ((0, 10000), (20, 20000)):
        fb = ao.melscale_fbanks(F, float(f0), float(f1), 512, 44100)
        ref = torchaudio.functional.melscale_fbanks(F, float(f0), float(f1), 512, 44100, None, "htk")
        assert torch.equal(fb, ref)
    fb = ao.melscale_fbanks(F, 0.0, 10000.0, 512, 44100)
    rows = (fb != 0).any(dim=1).nonzero().ravel()
    assert int(rows[0]) == 1 and int(rows[-1]) == 4000 and len(rows) == 4000
    assert int((fb != 0).sum()) == 7976


def test_stft_istft_match_torch():
    torch.manual_seed(0)
    x = torch.randn(2, 12000, dtype=torch.float64)
    win = torch.hann_window(W, dtype=torch.float64)
    ref = torch.stft(x, N, H, W, win, center=True, pad_mode="reflect", return_complex=True).numpy()
    got = ao.stft(x.numpy(), N, H, win.numpy())
    assert got.shape == ref.shape == (2, F, 1 + 12000 // H)
    assert np.abs(got - ref).max() < 1e-9
    spec = torch.randn(2, F, 24, dtype=torch.complex128)
    ref_i = torch.istft(spec, N, H, W, win).numpy()
    got_i = ao.istft(spec.numpy(), N, H, win.numpy())
    assert got_i.shape == ref_i.shape == (2, H * 23)
    assert np.abs(got_i - ref_i).max() < 1e-12
    with pytest.raises(ValueError):
        ao.stft(np.zeros((1, N // 2)), N, H, win.numpy())   # torch raises on pad >= length too


def test_griffinlim_matches_torchaudio_fp64_and_fp32():
    torch.manual_seed(1)
    T_ = 24
    mag = torch.rand(1, F, T_, dtype=torch.float64) * 10
    ang = torch.rand(1, F, T_, dtype=torch.complex128)
    gl = torchaudio.transforms.GriffinLim(n_fft=N, n_iter=5, win_length=W, hop_length=H, power=1.0,
                                          momentum=0.99, rand_init=True).double()
    ref = griffinlim_with_angles(gl, mag, ang).numpy()
    got = ao.griffinlim(mag.numpy(), N, H, ao.hann_window(W).double().numpy(), 5, 0.99, ang.numpy())
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-9
    # fp32 torchaudio vs fp64 oracle: bounded by fp32 rounding after 5 iterations
    gl32 = torchaudio.transforms.GriffinLim(n_fft=N, n_iter=5, win_length=W, hop_length=H, power=1.0,
                                            momentum=0.99, rand_init=True)
    ref32 = griffinlim_with_angles(gl32, mag.float(), ang.to(torch.complex64)).numpy()
    assert np.sqrt(np.mean((ref32 - got) ** 2)) / np.sqrt(np.mean(got ** 2)) < 2e-5


def test_inverse_mel_matches_torchaudio_lstsq():
    torch.manual_seed(2)
    conv = TorchaudioConverter()
    mel = torch.rand(2, 512, 8) * 1e6
    ref = conv.inverse_mel_scaler(mel).numpy()
    got = ao.inverse_mel(mel.numpy(), conv.mel_scaler.fb.numpy())
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 5e-6
    live = (conv.mel_scaler.fb != 0).any(dim=1).numpy()
    assert np.all(got[:, ~live] == 0)


def test_golden_torchaudio_vectors(golden):
    """the committed vectors are reproducible by the oracle"""
    g = golden["ta"]
    fb = ao.melscale_fbanks(F, 0.0, 10000.0, 512, 44100).numpy()
    lin = ao.inverse_mel(g["mel"], fb)
    assert np.linalg.norm(lin[:, g["live"]] - g["lin_live"]) / np.linalg.norm(g["lin_live"]) < 5e-6
    ang = np.zeros((1, F, 24), np.complex64)
    ang[:, g["live"]] = g["angles_live"]
    lin32 = np.zeros((1, F, 24), np.float32)
    lin32[:, g["live"]] = g["lin_live"]
    win = ao.hann_window(W).double().numpy()
    wave = ao.griffinlim(lin32, N, H, win, int(g["n_iter"]), 0.99, ang)
    ref = g["wave"]
    assert np.sqrt(np.mean((wave - ref) ** 2)) / np.sqrt(np.mean(ref ** 2)) < 2e-5
    mel_fwd = ao.mel_amplitudes_from_waveform(g["x"], fb, N, H, win)
    assert np.abs(mel_fwd - g["mel_fwd"]).max() / np.abs(g["mel_fwd"]).max() < 2e-6


def test_reference_fixture_known_answer(golden):
    """clip_2 WAV -> stereo PNG: the reference's own fixture pair, incl. EXIF MAX_VALUE."""
    g = golden["clip2"]
    wav = g["wav"].astype(np.float32).T            # (2, 250400) raw int16-valued floats
    assert int(g["rate"]) == 44100 and wav.shape == (2, 250400)
    fb = ao.melscale_fbanks(F, 0.0, 10000.0, 512, 44100).numpy()
    mel = ao.mel_amplitudes_from_waveform(wav, fb, N, H, ao.hann_window(W).double().numpy())
    mel = mel.astype(np.float32)
    assert mel.shape == (2, 512, 568)
    exif = dict(zip(g["exif_keys"].tolist(), g["exif_stereo"].tolist()))
    assert abs(float(mel.max()) - exif[11080]) <= 8.0          # MAX_VALUE = 46801012.0 (fp32 ulp = 4)
    img, mx = ao.image_array_from_spectrogram(mel, power=0.25)
    png = g["stereo_png"]
    assert img.shape == png.shape == (512, 568, 3)
    diff = np.abs(img.astype(np.int16) - png.astype(np.int16))
    assert diff.max() <= 1 and (diff != 0).mean() < 5e-4
    assert np.all(png[:, :, 0] == 0)                           # stereo => R plane zero
    # the mono fixture is the left channel (G plane) of the stereo image
    assert np.array_equal(golden["clip2"]["mono_png"][:, :, 0], png[:, :, 1])


def test_image_roundtrip_quantisation(golden):
    rgb = golden["og_beat"]["rgb"]
    assert rgb.shape == (512, 512, 3) and str(golden["og_beat"]["mode"]) == "P"
    spec = ao.spectrogram_from_image_array(rgb, power=0.25, stereo=False, max_value=30e6)
    assert spec.shape == (1, 512, 512) and spec.dtype == np.float32
    assert 29.0 < spec.min() < 29.2 and 2.7e7 < spec.max() < 2.8e7
    img, mx = ao.image_array_from_spectrogram(spec, power=0.25)
    # mirrors test/image_util_test.py:53-65 tolerances on the round trip
    back = ao.spectrogram_from_image_array(img, power=0.25, stereo=False, max_value=mx)
    assert back.shape == spec.shape
    assert np.isclose(back.max(), spec.max())
    assert np.allclose(back, spec, rtol=0.15, atol=spec.max() * 1e-4)


def test_int16_truncation_and_joint_normalisation():
    w = np.array([[0.5, -0.25, 0.999], [1.0, -2.0, 0.1]], np.float32)
    pcm = ao.int16_from_waveform(w, normalize=True)
    assert pcm.shape == (3, 2) and pcm.dtype == np.int16
    assert pcm[1, 1] == -32767 and pcm[0, 0] == int(0.5 * np.float32(32767 / 2.0))


def test_host_slerp_bit_exact_against_reference_vectors():
    """riffusion.util.torch_util.slerp (host mode) reproduces the reference's own function bit for bit on fp16 and fp32
    tensors, including the lerp branch for nearly parallel vectors (tests/golden/make_golden_host.py); the oracle's
    slerp agrees with it"""
    import torch

    from oracle import unet_oracle as uo
    from riffusion.util import torch_util

    from pathlib import Path

    g = np.load(Path(__file__).resolve().parent / "golden" / "host_vectors.npz")
    for name in ("f16", "f32", "par16"):
        a, b = torch.from_numpy(g[f"slerp_{name}_a"]), torch.from_numpy(g[f"slerp_{name}_b"])
        for i, t in enumerate(g["ts"]):
            got = torch_util.slerp(float(t), a, b)
            assert got.dtype == a.dtype
            assert np.array_equal(got.numpy(), g[f"slerp_{name}_out"][i]), (name, t)
            ref32 = uo.slerp(float(t), a.float(), b.float()).numpy()
            assert np.allclose(got.float().numpy(), ref32, atol=4e-3 if a.dtype == torch.float16 else 1e-6)
