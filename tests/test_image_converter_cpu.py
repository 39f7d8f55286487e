"""SpectrogramImageConverter's host logic (channel policy, EXIF tags, argument plumbing) with the GPU-backed
SpectrogramConverter replaced by a recorder — reference behaviour: riffusion/spectrogram_image_converter.py:23-91."""
import numpy as np
import pytest

from riffusion import spectrogram_image_converter as sic
from riffusion.spectrogram_params import SpectrogramParams
from riffusion.util import image_util


class _Seg:
    def __init__(self, channels, frame_rate=44100):
        self.channels, self.frame_rate, self.history = channels, frame_rate, []

    def set_channels(self, n):
        s = _Seg(n, self.frame_rate)
        s.history = self.history + [n]
        return s


class _FakeConverter:
    def __init__(self, params, device):
        self.params, self.device, self.seen = params, device, []

    def spectrogram_from_audio(self, segment):
        self.seen.append(segment)
        c = 2 if self.params.stereo else 1
        rng = np.random.default_rng(0)
        return (rng.random((c, 512, 40)).astype(np.float32) ** 4) * 4.2e7

    def audio_from_spectrogram(self, spectrogram, apply_filters=True):
        return ("AUDIO", spectrogram, apply_filters)


@pytest.fixture()
def patched(monkeypatch):
    monkeypatch.setattr(sic, "SpectrogramConverter", _FakeConverter)


@pytest.mark.parametrize("stereo,channels,expect_channels,msg", [
    (True, 1, 2, "cloning channel"), (True, 6, 2, "reducing to stereo"), (True, 2, 2, None),
    (False, 2, 1, "setting to mono"), (False, 1, 1, None)])
def test_channel_policy_and_exif(patched, capsys, stereo, channels, expect_channels, msg):
    p = SpectrogramParams(stereo=stereo)
    conv = sic.SpectrogramImageConverter(p, device="cuda")
    assert conv.p is p and conv.device == "cuda" and isinstance(conv.converter, _FakeConverter)
    img = conv.spectrogram_image_from_audio(_Seg(channels))
    out = capsys.readouterr().out
    assert (msg in out) if msg else ("WARNING" not in out)
    assert conv.converter.seen[-1].channels == expect_channels
    spec = conv.converter.spectrogram_from_audio(None)
    exif = image_util.exif_from_image(img)
    assert exif[SpectrogramParams.ExifTags.MAX_VALUE.name] == pytest.approx(float(spec.max()))
    assert SpectrogramParams.from_exif(img.getexif()) == p
    ref = image_util.image_from_spectrogram(spec, power=p.power_for_image)
    assert np.array_equal(np.array(img), np.array(ref))


def test_sample_rate_mismatch_and_reverse_path(patched):
    conv = sic.SpectrogramImageConverter(SpectrogramParams(), device="cuda")
    with pytest.raises(AssertionError, match="Sample rate mismatch"):
        conv.spectrogram_image_from_audio(_Seg(1, frame_rate=22050))
    img = conv.spectrogram_image_from_audio(_Seg(1))
    tag, spec, filt = conv.audio_from_spectrogram_image(img, apply_filters=False, max_value=1e6)
    assert tag == "AUDIO" and filt is False
    assert np.array_equal(spec, image_util.spectrogram_from_image(img, max_value=1e6, power=0.25, stereo=False))
