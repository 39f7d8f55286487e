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
