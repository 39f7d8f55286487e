"""GPU parity tests for hot path (a): every check goes through the reference-facing Python API
(SpectrogramConverter and its transform attributes), i.e. through the C-ABI of librf_b200.so, and
compares with (i) the installed torchaudio transforms built with the reference's arguments
(oracle/torchaudio_ref.py), (ii) the fp64 oracle restatement (oracle/audio_oracle.py) and (iii) the
committed golden fixtures.  Nothing here reads /root/reference.

Tolerances
  * STFT / mel / inverse mel (single linear passes): max-abs error <= 2e-6 of the output peak.
  * Griffin-Lim waveform: RMS error on the peak-normalised waveform <= 1e-4 against
    torchaudio with identical injected initial phases.  Griffin-Lim amplifies fp32 rounding (two
    correct fp32 FFTs drift apart ~1e-4 relative over 32 iterations, SURVEY §7 hard part 2), so the
    fp64 oracle is the tie-breaker: our distance to the exact recurrence must not exceed
    max(2x torchaudio's own distance, 2e-5).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, W, H, F = 17640, 4410, 441, 8821


@pytest.fixture(scope="module")
def conv(native_lib):
    from riffusion.spectrogram_converter import SpectrogramConverter
    from riffusion.spectrogram_params import SpectrogramParams

    assert torch.cuda.is_available()
    return SpectrogramConverter(SpectrogramParams(), device="cuda")


@pytest.fixture(scope="module")
def ta():
    from oracle.torchaudio_ref import TorchaudioConverter

    return TorchaudioConverter()


def _norm_rms(a: torch.Tensor, ref: torch.Tensor) -> float:
    peak = ref.abs().amax(dim=-1, keepdim=True)
    return float((((a - ref) / peak) ** 2).mean().sqrt())


def test_extension_is_loaded_and_on_device(conv):
    from riffusion import _native

    assert "librf_b200.so" in str(_native.lib()._name)
    with pytest.raises(_native.NativeError):
        conv.mel_amplitudes_from_waveform(torch.zeros(1, 20000))      # CPU tensor: no fallback


@pytest.mark.parametrize("L", [H * 24, 30000, 250400])     # exact frames / ragged / the fixture clip length
def test_stft_matches_torch(conv, ta, L):
    torch.manual_seed(L)
    x = torch.randn(2, L) * 3000
    ref = ta.spectrogram_func(x)
    got = conv.spectrogram_func(x.cuda()).cpu()
    assert got.shape == ref.shape == (2, F, 1 + L // H)
    assert float((got - ref).abs().max() / ref.abs().max()) < 2e-6


def test_stft_too_short_raises_like_torch(conv):
    with pytest.raises(ValueError, match="Padding size should be less"):
        conv.spectrogram_func(torch.zeros(1, N // 2, device="cuda"))
    with pytest.raises(ValueError, match="Padding size should be less"):
        conv.waveform_from_mel_amplitudes(torch.ones(1, 512, 21, device="cuda"))   # hop*20 <= n_fft/2


def test_mel_and_inverse_mel_match_torchaudio(conv, ta):
    torch.manual_seed(3)
    spec = torch.rand(2, F, 40) * 1e5
    ref = ta.mel_scaler(spec)
    got = conv.mel_scaler(spec.cuda()).cpu()
    assert float((got - ref).abs().max() / ref.abs().max()) < 2e-6
    mel = (torch.rand(3, 512, 70) ** 4) * 3e7
    ref_i = ta.inverse_mel_scaler(mel)
    got_i = conv.inverse_mel_scaler(mel.cuda()).cpu()
    assert got_i.shape == ref_i.shape == (3, F, 70)
    assert float((got_i - ref_i).norm() / ref_i.norm()) < 5e-6
    live = (ta.mel_scaler.fb != 0).any(dim=1)
    assert torch.all(got_i[:, ~live] == 0) and torch.all(got_i >= 0)
    with pytest.raises(ValueError, match="Expected an input with 512 mel bins"):
        conv.inverse_mel_scaler(torch.zeros(1, 256, 30, device="cuda"))


def test_forward_path_fixture_known_answer(conv, golden):
    """reference fixture: clip_2 WAV -> stereo PNG incl. EXIF MAX_VALUE, through the fused kernel and
    the device image quantiser"""
    from riffusion.util import image_util

    g = golden["clip2"]
    wav = torch.from_numpy(g["wav"].astype(np.float32).T.copy())
    mel = conv.mel_amplitudes_from_waveform(wav.cuda())
    assert mel.shape == (2, 512, 568)                     # width == round(5678/10) (audio_to_image_test.py:73)
    exif = dict(zip(g["exif_keys"].tolist(), g["exif_stereo"].tolist()))
    img, mx = image_util.image_from_spectrogram_device(mel, power=0.25)
    assert abs(float(mx) - exif[11080]) <= 16.0
    png = g["stereo_png"]
    diff = (img.cpu().numpy().astype(np.int16) - png.astype(np.int16))
    assert np.abs(diff).max() <= 1 and (diff != 0).mean() < 1e-3
    # host quantiser on the same amplitudes agrees with the device one except at truncation boundaries
    host = np.array(image_util.image_from_spectrogram(mel.cpu().numpy(), power=0.25))
    d2 = np.abs(host.astype(np.int16) - img.cpu().numpy().astype(np.int16))
    assert d2.max() <= 1 and (d2 != 0).mean() < 1e-3


def test_image_to_mel_device_matches_host(conv, golden):
    from PIL import Image

    from riffusion.util import image_util

    rgb = golden["og_beat"]["rgb"]
    pil = Image.fromarray(rgb, mode="RGB")
    for stereo in (False, True):
        host = image_util.spectrogram_from_image(pil, power=0.25, stereo=stereo, max_value=30e6)
        dev = image_util.spectrogram_from_image_device(pil, power=0.25, stereo=stereo, max_value=30e6).cpu().numpy()
        assert dev.shape == host.shape
        assert np.abs(dev - host).max() <= 4e-6 * host.max()          # powf vs numpy power: <= 2 ulp-ish


def test_wave_to_int16_matches_numpy(conv):
    from riffusion import _native
    from oracle import audio_oracle as ao

    torch.manual_seed(5)
    w = torch.randn(2, 50001) * 0.3
    ref = ao.int16_from_waveform(w.numpy(), normalize=True)
    wd = w.cuda()
    pcm = torch.empty((50001, 2), dtype=torch.int16, device="cuda")
    scratch = torch.zeros(1, dtype=torch.float32, device="cuda")
    _native.check(_native.lib().rf_wave_to_int16(wd.data_ptr(), 2, 50001, 1, pcm.data_ptr(), scratch.data_ptr(),
                                                 _native.stream_ptr(wd.device)))
    d = np.abs(pcm.cpu().numpy().astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3


def test_golden_torchaudio_vectors(conv, golden):
    g = golden["ta"]
    live = torch.from_numpy(g["live"])
    lin = conv.inverse_mel_scaler(torch.from_numpy(g["mel"]).cuda()).cpu()
    assert float((lin[:, live] - torch.from_numpy(g["lin_live"])).norm() / np.linalg.norm(g["lin_live"])) < 5e-6
    from riffusion.spectrogram_converter import SpectrogramConverter
    from riffusion.spectrogram_params import SpectrogramParams

    c4 = SpectrogramConverter(SpectrogramParams(num_griffin_lim_iters=int(g["n_iter"])), device="cuda")
    ang = torch.zeros(1, F, 24, dtype=torch.complex64)
    ang[:, live] = torch.from_numpy(g["angles_live"])
    wave = c4.waveform_from_mel_amplitudes(torch.from_numpy(g["mel"]).cuda(), ang.cuda()).cpu()
    assert _norm_rms(wave, torch.from_numpy(g["wave"])) < 1e-5
    mel_fwd = conv.mel_amplitudes_from_waveform(torch.from_numpy(g["x"]).cuda()).cpu()
    assert float((mel_fwd - torch.from_numpy(g["mel_fwd"])).abs().max() / g["mel_fwd"].max()) < 2e-6


@pytest.mark.parametrize("T_,n_iter,B", [(24, 1, 2), (35, 4, 1), (64, 8, 3)])
def test_griffinlim_small_vs_torchaudio_and_fp64(conv, ta, T_, n_iter, B):
    """generic GriffinLim callable (full band) and the fused pruned path on the same input"""
    from oracle import audio_oracle as ao
    from oracle.torchaudio_ref import griffinlim_with_angles
    from riffusion.spectrogram_converter import SpectrogramConverter
    from riffusion.spectrogram_params import SpectrogramParams
    import torchaudio

    torch.manual_seed(100 * T_ + n_iter)
    mel = (torch.rand(B, 512, T_) ** 4) * 3e7
    ang = torch.rand(B, F, T_, dtype=torch.complex64)
    lin = ta.inverse_mel_scaler(mel)
    gl = torchaudio.transforms.GriffinLim(n_fft=N, n_iter=n_iter, win_length=W, hop_length=H, power=1.0,
                                          momentum=0.99, rand_init=True)
    ref = griffinlim_with_angles(gl, lin, ang)
    o64 = torch.from_numpy(ao.griffinlim(lin.numpy(), N, H, ao.hann_window(W).double().numpy(), n_iter, 0.99,
                                         ang.numpy())).float()
    c = SpectrogramConverter(SpectrogramParams(num_griffin_lim_iters=n_iter), device="cuda")
    full = c.inverse_spectrogram_func.forward(lin.cuda(), ang.cuda()).cpu()
    fused = c.waveform_from_mel_amplitudes(mel.cuda(), ang.cuda()).cpu()
    assert full.shape == fused.shape == ref.shape == (B, H * (T_ - 1))
    err_ta = _norm_rms(ref, o64)
    for got in (full, fused):
        assert _norm_rms(got, ref) < 1e-4
        assert _norm_rms(got, o64) <= max(2 * err_ta, 2e-5)


def test_griffinlim_og_beat_32_iters_vs_torchaudio(conv, ta, golden):
    """BASELINE config 1 input (seed_images/og_beat.png amplitudes, 512x512, 32 iterations):
    <= 1e-4 RMS on the peak-normalised waveform against torchaudio on identical initial phases."""
    from oracle import audio_oracle as ao

    spec = ao.spectrogram_from_image_array(golden["og_beat"]["rgb"], power=0.25, stereo=False, max_value=30e6)
    mel = torch.from_numpy(spec)
    torch.manual_seed(0)
    ang = torch.rand(1, F, 512, dtype=torch.complex64)
    ref = ta.waveform_from_mel_amplitudes(mel, ang)
    got = conv.waveform_from_mel_amplitudes(mel.cuda(), ang.cuda()).cpu()
    assert got.shape == ref.shape == (1, 225351)
    rms = _norm_rms(got, ref)
    print(f"og_beat 32-iter normalised RMS vs torchaudio: {rms:.3e}")
    assert rms < 1e-4
    # duration / sample-rate contract of test/image_to_audio_test.py:55-60 (5.11 s at 44.1 kHz)
    assert abs(got.shape[1] / 44100 - 5.11) < 0.01


@pytest.mark.parametrize("T_,n_iter,B", [(64, 4, 1), (131, 6, 2), (512, 8, 1)])
def test_griffinlim_decimated_loop_vs_full_rate(conv, ta, T_, n_iter, B):
    """The half-rate inner loop (odd samples + full-rate edge strips built from both sample parities) is a re-association of the same
    arithmetic: against the full-rate loop of the same library, the fp64 oracle and torchaudio."""
    from oracle import audio_oracle as ao
    from riffusion.spectrogram_converter import SpectrogramConverter, get_plan
    from riffusion.spectrogram_params import SpectrogramParams

    torch.manual_seed(7 * T_ + n_iter)
    mel = (torch.rand(B, 512, T_) ** 4) * 3e7
    ang = torch.rand(B, F, T_, dtype=torch.complex64)
    prm = SpectrogramParams(num_griffin_lim_iters=n_iter)
    c = SpectrogramConverter(prm, device="cuda")
    plan = get_plan(prm, full_band=False)
    try:
        assert plan.set_decimation(False) is False
        full = c.waveform_from_mel_amplitudes(mel.cuda(), ang.cuda()).cpu()
        assert plan.set_decimation(True) is True
        dec = c.waveform_from_mel_amplitudes(mel.cuda(), ang.cuda()).cpu()
    finally:
        plan.set_decimation(True)
    assert not torch.equal(full, dec)
    import torchaudio

    from oracle.torchaudio_ref import griffinlim_with_angles

    lin = ta.inverse_mel_scaler(mel)
    gl = torchaudio.transforms.GriffinLim(n_fft=N, n_iter=n_iter, win_length=W, hop_length=H, power=1.0,
                                          momentum=0.99, rand_init=True)
    ref = griffinlim_with_angles(gl, lin, ang)
    o64 = torch.from_numpy(ao.griffinlim(lin.numpy(), N, H, ao.hann_window(W).double().numpy(), n_iter, 0.99,
                                         ang.numpy())).float()
    err_ta, err_full, err_dec = _norm_rms(ref, o64), _norm_rms(full, o64), _norm_rms(dec, o64)
    print(f"T={T_} iters={n_iter}: normalised RMS vs fp64 - torchaudio {err_ta:.2e}, full-rate {err_full:.2e}, "
          f"decimated {err_dec:.2e}; decimated vs full-rate {_norm_rms(dec, full):.2e}")
    assert err_full <= max(2 * err_ta, 2e-5)
    assert err_dec <= max(3 * err_ta, 2e-5)
    assert _norm_rms(dec, ref) < 1e-4


def test_griffinlim_full_size_properties(conv):
    """size-independent properties at BASELINE's full size (batch 16 x 512 frames): batch
    independence (clip b of a batch == the same clip alone, bit-exact), determinism, finiteness,
    and Griffin-Lim's defining property — spectral convergence improves with iterations."""
    from riffusion.spectrogram_converter import SpectrogramConverter
    from riffusion.spectrogram_params import SpectrogramParams

    torch.manual_seed(11)
    B, T_ = 16, 512
    mel = (torch.rand(B, 512, T_, device="cuda") ** 4) * 3e7
    ang = torch.rand(B, F, T_, dtype=torch.complex64, device="cuda")
    w1 = conv.waveform_from_mel_amplitudes(mel, ang)
    w2 = conv.waveform_from_mel_amplitudes(mel, ang)
    assert torch.equal(w1, w2)                                        # deterministic (no atomics)
    solo = conv.waveform_from_mel_amplitudes(mel[5:6].contiguous(), ang[5:6].contiguous())
    assert torch.equal(solo[0], w1[5])                                # clips are independent
    assert torch.isfinite(w1).all()
    lin = conv.inverse_mel_scaler(mel[:2].contiguous())

    def inconsistency(n_iter):
        c = SpectrogramConverter(SpectrogramParams(num_griffin_lim_iters=n_iter), device="cuda")
        w = c.waveform_from_mel_amplitudes(mel[:2].contiguous(), ang[:2].contiguous())
        mag = c.spectrogram_func(w).abs()
        return float((mag - lin).norm() / lin.norm())

    e0, e8, e32 = inconsistency(0), inconsistency(8), inconsistency(32)
    assert e32 < e8 < e0


def test_ragged_and_odd_frame_counts(conv, ta):
    """odd T (dangling half pair), T not a multiple of the 16-frame overlap-add chunk"""
    from riffusion.spectrogram_converter import SpectrogramConverter
    from riffusion.spectrogram_params import SpectrogramParams

    c = SpectrogramConverter(SpectrogramParams(num_griffin_lim_iters=2), device="cuda")
    for T_ in (23, 33, 49):
        torch.manual_seed(T_)
        mel = (torch.rand(1, 512, T_) ** 4) * 3e7
        ang = torch.rand(1, F, T_, dtype=torch.complex64)
        import torchaudio
        from oracle.torchaudio_ref import griffinlim_with_angles

        gl = torchaudio.transforms.GriffinLim(n_fft=N, n_iter=2, win_length=W, hop_length=H, power=1.0,
                                              momentum=0.99, rand_init=True)
        ref = griffinlim_with_angles(gl, ta.inverse_mel_scaler(mel), ang)
        got = c.waveform_from_mel_amplitudes(mel.cuda(), ang.cuda()).cpu()
        assert _norm_rms(got, ref) < 1e-5


def test_wide_band_params_like_reference_tests(native_lib):
    """the reference's converter tests use f in [20, 20000] (spectrogram_converter_test.py:46-53)"""
    from oracle.torchaudio_ref import TorchaudioConverter
    from riffusion.spectrogram_converter import SpectrogramConverter
    from riffusion.spectrogram_params import SpectrogramParams

    p = SpectrogramParams(min_frequency=20, max_frequency=20000, num_griffin_lim_iters=2)
    c = SpectrogramConverter(p, device="cuda")
    t = TorchaudioConverter(f_min=20, f_max=20000, n_iter=2)
    torch.manual_seed(9)
    x = torch.randn(1, 20000) * 3000
    ref = t.mel_amplitudes_from_waveform(x)
    got = c.mel_amplitudes_from_waveform(x.cuda()).cpu()
    assert float((got - ref).abs().max() / ref.abs().max()) < 2e-6
    mel = (torch.rand(1, 512, 30) ** 4) * 3e7
    ang = torch.rand(1, F, 30, dtype=torch.complex64)
    refw = t.waveform_from_mel_amplitudes(mel, ang)
    gotw = c.waveform_from_mel_amplitudes(mel.cuda(), ang.cuda()).cpu()
    assert _norm_rms(gotw, refw) < 1e-5


def test_random_init_uses_torch_global_rng(conv):
    """rand_init=True draws torch.rand(cfloat) from the global generator (functional.py:309-310):
    same seed => same waveform, different seed => different waveform"""
    mel = (torch.rand(1, 512, 40, device="cuda") ** 4) * 3e7
    torch.manual_seed(7)
    a = conv.waveform_from_mel_amplitudes(mel)
    torch.manual_seed(7)
    b = conv.waveform_from_mel_amplitudes(mel)
    torch.manual_seed(8)
    c = conv.waveform_from_mel_amplitudes(mel)
    assert torch.equal(a, b) and not torch.equal(a, c)


@pytest.mark.parametrize("sr,kw", [(48000, {}), (22050, {}), (44100, dict(window_duration_ms=50, padded_duration_ms=200, step_size_ms=5, num_frequencies=256))])
def test_generic_engine_other_geometries_vs_torchaudio(native_lib, sr, kw):
    """sample rates / window settings the prime-factor engine does not cover (cli.py:40-52 builds the params from the
    file's frame rate): 48 kHz (n_fft 19200 = 2 * 2^7 3 5^2), 22.05 kHz (win 2205 odd, hop 220 does not divide it: the
    overlap-add envelope is not constant), and non-default window / padding / step durations.
    STFT, STFT+mel, inverse mel and Griffin-Lim against the installed torchaudio transforms with the reference's arguments."""
    import torchaudio

    from oracle.torchaudio_ref import TorchaudioConverter, griffinlim_with_angles
    from riffusion.spectrogram_converter import SpectrogramConverter
    from riffusion.spectrogram_params import SpectrogramParams

    n_iter = 4
    p = SpectrogramParams(sample_rate=sr, num_griffin_lim_iters=n_iter, **kw)
    c = SpectrogramConverter(p, device="cuda")
    t = TorchaudioConverter(sample_rate=sr, n_fft=p.n_fft, win_length=p.win_length, hop_length=p.hop_length, n_mels=p.num_frequencies,
                            n_iter=n_iter)
    Fg = p.n_fft // 2 + 1
    torch.manual_seed(sr)
    x = torch.randn(2, p.n_fft + 37 * p.hop_length + 11) * 3000
    ref = t.spectrogram_func(x)
    got = c.spectrogram_func(x.cuda()).cpu()
    assert got.shape == ref.shape == (2, Fg, 1 + x.shape[1] // p.hop_length)
    assert float((got - ref).abs().max() / ref.abs().max()) < 3e-6
    mel_ref = t.mel_amplitudes_from_waveform(x)
    mel_got = c.mel_amplitudes_from_waveform(x.cuda()).cpu()
    assert float((mel_got - mel_ref).abs().max() / mel_ref.abs().max()) < 3e-6
    T_ = 60
    mel = (torch.rand(2, p.num_frequencies, T_) ** 4) * 3e7
    lin_ref = t.inverse_mel_scaler(mel)
    lin_got = c.inverse_mel_scaler(mel.cuda()).cpu()
    assert float((lin_got - lin_ref).norm() / lin_ref.norm()) < 1e-5
    ang = torch.rand(2, Fg, T_, dtype=torch.complex64)
    gl = torchaudio.transforms.GriffinLim(n_fft=p.n_fft, n_iter=n_iter, win_length=p.win_length, hop_length=p.hop_length, power=1.0,
                                          momentum=0.99, rand_init=True)
    wave_ref = griffinlim_with_angles(gl, lin_ref, ang)
    wave_full = c.inverse_spectrogram_func.forward(lin_ref.cuda(), ang.cuda()).cpu()
    wave_fused = c.waveform_from_mel_amplitudes(mel.cuda(), ang.cuda()).cpu()
    assert wave_full.shape == wave_fused.shape == wave_ref.shape == (2, p.hop_length * (T_ - 1))
    assert _norm_rms(wave_full, wave_ref) < 2e-5 and _norm_rms(wave_fused, wave_ref) < 5e-5
