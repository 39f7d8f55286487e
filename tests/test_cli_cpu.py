"""CLI surface (reference: riffusion/cli.py): same six commands, same flags; print-exif and sample-clips run on
CPU here (mirrors test/print_exif_test.py and test/sample_clips_test.py); the GPU commands are exercised in
tests/test_cli_gpu.py."""
import io
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest
from PIL import Image


def test_parser_has_reference_commands_and_flags():
    from riffusion import cli

    parser = cli.build_parser()
    sub = next(a for a in parser._actions if a.dest == "command")
    assert set(sub.choices) == {"audio-to-image", "image-to-audio", "sample-clips", "print-exif",
                                "audio-to-images-batch", "sample-clips-batch"}
    a2i = {o for act in sub.choices["audio-to-image"]._actions for o in act.option_strings}
    assert {"--audio", "--image", "--step-size-ms", "--num-frequencies", "--min-frequency", "--max-frequency",
            "--window-duration-ms", "--padded-duration-ms", "--power-for-image", "--stereo", "--device"} <= a2i
    ns = parser.parse_args(["audio-to-image", "--audio", "a.wav", "--image", "b.png", "--stereo", "--max-frequency", "20000"])
    assert ns.stereo is True and ns.max_frequency == 20000 and ns.device == "cuda"


def test_print_exif_and_params_roundtrip(tmp_path):
    from riffusion import cli
    from riffusion.spectrogram_params import SpectrogramParams

    p = SpectrogramParams(stereo=True, max_frequency=20000, min_frequency=20)
    img = Image.fromarray(np.zeros((8, 8, 3), np.uint8))
    exif = img.getexif()
    tags = p.to_exif()
    tags[SpectrogramParams.ExifTags.MAX_VALUE.value] = 46801012.0
    exif.update(tags.items())
    path = tmp_path / "x.png"
    img.save(path, exif=exif, format="PNG")
    loaded = Image.open(path)
    assert SpectrogramParams.from_exif(loaded.getexif()) == p            # audio_to_image_test.py:89-99
    buf = io.StringIO()
    with redirect_stdout(buf):
        cli.main(["print-exif", "--image", str(path)])
    out = buf.getvalue()
    assert "NUM_FREQUENCIES      =             512" in out and "SAMPLE_RATE          =           44100" in out   # print_exif_test.py:31-32


def test_sample_clips_wav(tmp_path):
    from scipy.io import wavfile

    from riffusion import cli
    from riffusion.util.audio_util import AudioSegment

    rate = 44100
    t = np.arange(rate * 3) / rate
    stereo = np.stack([np.sin(2 * np.pi * 440 * t), np.sin(2 * np.pi * 220 * t)], axis=1)
    wavfile.write(tmp_path / "in.wav", rate, (stereo * 20000).astype(np.int16))
    cli.main(["sample-clips", "--audio", str(tmp_path / "in.wav"), "--output-dir", str(tmp_path / "out"),
              "--num-clips", "3", "--duration-ms", "500", "--mono", "--seed", "7"])
    clips = sorted((tmp_path / "out").glob("clip_*_ms_duration_500_ms.wav"))
    assert len(clips) == 3                                               # sample_clips_test.py: count / extension
    for c in clips:
        seg = AudioSegment.from_file(str(c))
        assert seg.channels == 1 and seg.frame_rate == rate and abs(len(seg) - 500) <= 1


class _FakeImageConverter:
    """stands in for the GPU-backed SpectrogramImageConverter: records what the CLI passes to it"""
    instances = []

    def __init__(self, params, device):
        self.p, self.device, self.calls = params, device, []
        _FakeImageConverter.instances.append(self)

    def spectrogram_image_from_audio(self, segment):
        from PIL import Image

        from riffusion.spectrogram_params import SpectrogramParams

        self.calls.append(("to_image", segment.channels, segment.frame_rate))
        img = Image.new("RGB", (64, self.p.num_frequencies), (1, 2, 3))
        tags = self.p.to_exif()
        tags[SpectrogramParams.ExifTags.MAX_VALUE.value] = 123.0
        img.getexif().update(tags.items())
        return img

    def audio_from_spectrogram_image(self, image, apply_filters=True, max_value=30e6):
        import numpy as np

        from riffusion.util.audio_util import AudioSegment

        self.calls.append(("to_audio", image.size))
        n = self.p.sample_rate // 10
        return AudioSegment((np.zeros((n, 2 if self.p.stereo else 1)) + 100).astype(np.int16), self.p.sample_rate)


def test_commands_with_recorded_converter(tmp_path, monkeypatch, capsys):
    """audio-to-image / image-to-audio / audio-to-images-batch argument plumbing (flags -> SpectrogramParams, EXIF written
    and read back, channel / sample-rate conforming in the batch command, unreadable files skipped) without a GPU"""
    import numpy as np
    from PIL import Image
    from scipy.io import wavfile

    from riffusion import cli
    from riffusion.spectrogram_params import SpectrogramParams

    monkeypatch.setattr(cli, "SpectrogramImageConverter", _FakeImageConverter)
    _FakeImageConverter.instances.clear()
    rate = 22050
    wav = (np.sin(np.arange(rate) / 20.0) * 8000).astype(np.int16)
    wavfile.write(tmp_path / "a.wav", rate, wav)
    cli.main(["audio-to-image", "--audio", str(tmp_path / "a.wav"), "--image", str(tmp_path / "a.png"), "--stereo",
              "--num-frequencies", "256", "--max-frequency", "8000", "--power-for-image", "0.5", "--device", "cuda:1"])
    conv = _FakeImageConverter.instances[-1]
    assert conv.device == "cuda:1"
    assert conv.p == SpectrogramParams(sample_rate=rate, stereo=True, num_frequencies=256, max_frequency=8000,
                                       power_for_image=0.5)
    assert conv.calls == [("to_image", 1, rate)]
    img = Image.open(tmp_path / "a.png")
    assert img.format == "PNG" and SpectrogramParams.from_exif(img.getexif()) == conv.p     # EXIF survives the save
    assert f"Wrote {tmp_path / 'a.png'}" in capsys.readouterr().out

    cli.main(["image-to-audio", "--image", str(tmp_path / "a.png"), "--audio", str(tmp_path / "back.wav")])
    conv2 = _FakeImageConverter.instances[-1]
    assert conv2.p == conv.p and conv2.calls == [("to_audio", (64, 256))] and conv2.device == "cuda"
    r2, back = wavfile.read(tmp_path / "back.wav")
    assert r2 == rate and back.shape == (rate // 10, 2)
    assert "seconds)" in capsys.readouterr().out
    # an image without our EXIF tags falls back to the defaults with the reference's warning (cli.py:79-83)
    Image.new("RGB", (32, 512)).save(tmp_path / "plain.png")
    cli.main(["image-to-audio", "--image", str(tmp_path / "plain.png"), "--audio", str(tmp_path / "plain.wav")])
    assert "Using defaults" in capsys.readouterr().out and _FakeImageConverter.instances[-1].p == SpectrogramParams()

    # batch: stereo default -> mono files are widened, junk files are skipped (resampling needs pydub: same rate here)
    clips = tmp_path / "clips"
    clips.mkdir()
    wavfile.write(clips / "x.wav", rate, wav)
    wavfile.write(clips / "y.wav", rate, np.stack([wav, wav], axis=1))
    (clips / "notes.txt").write_text("not audio")
    cli.main(["audio-to-images-batch", "--audio-dir", str(clips), "--output-dir", str(tmp_path / "imgs"),
              "--image-extension", "png", "--num-threads", "2", "--sample-rate", str(rate)])
    shared = _FakeImageConverter.instances[-1]
    assert sorted(shared.calls) == [("to_image", 2, rate), ("to_image", 2, rate)] and shared.p.stereo
    assert sorted(p.name for p in (tmp_path / "imgs").iterdir()) == ["x.png", "y.png"]
