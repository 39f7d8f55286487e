"""CLI end to end on the GPU, mirroring the reference's CLI tests on its own fixture (test/audio_to_image_test.py,
test/image_to_audio_test.py, test/spectrogram_image_converter_test.py): clip_2 WAV -> PNG -> WAV."""
import numpy as np
import pytest
from PIL import Image

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("stereo", [False, True])
def test_audio_to_image_then_image_to_audio(native_lib, golden, tmp_path, stereo):
    from scipy.io import wavfile

    from riffusion import cli
    from riffusion.spectrogram_params import SpectrogramParams
    from riffusion.util.audio_util import AudioSegment

    g = golden["clip2"]
    wavfile.write(tmp_path / "clip.wav", int(g["rate"]), g["wav"])
    args = ["audio-to-image", "--audio", str(tmp_path / "clip.wav"), "--image", str(tmp_path / "clip.png")]
    cli.main(args + (["--stereo"] if stereo else []))
    img = Image.open(tmp_path / "clip.png")
    assert img.mode == "RGB" and img.width == round(5678 / 10) and img.height == 512      # audio_to_image_test.py:73-75
    arr = np.array(img)
    if stereo:
        assert np.all(arr[:, :, 0] == 0)                                                  # :81-83
        diff = np.abs(arr.astype(np.int16) - g["stereo_png"].astype(np.int16))
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-3                              # the reference's own fixture image
    else:
        assert np.array_equal(arr[:, :, 0], arr[:, :, 1]) and np.array_equal(arr[:, :, 0], arr[:, :, 2])   # :84-87
    params = SpectrogramParams.from_exif(img.getexif())
    assert params == SpectrogramParams(stereo=stereo)                                     # :89-99
    cli.main(["image-to-audio", "--image", str(tmp_path / "clip.png"), "--audio", str(tmp_path / "out.wav")])
    seg = AudioSegment.from_file(str(tmp_path / "out.wav"))
    assert seg.frame_rate == 44100                                                        # image_to_audio_test.py:55
    assert abs(seg.duration_seconds * 1000 - 5678) < 12                                   # :58-60 (10 ms + 1 hop)
    assert seg.channels == (2 if stereo else 1) and seg.sample_width == 2                 # :63-67
