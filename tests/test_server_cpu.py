"""(f)-4 response contract: `riffusion.server.compute_request` / `run_inference` (reference: riffusion/server.py:66-183) with
a recorded pipeline and converter — request validation, seed / mask lookup, error tuples, and the JSON fields
(`data:image/jpeg;base64,` / `data:audio/...;base64,` / duration_s).  No GPU: the heavy parts are stand-ins here; the GPU
version is tests/test_cli_gpu.py::test_server_compute_request_on_gpu."""
import base64
import io
import json
from pathlib import Path

import numpy as np
from PIL import Image

GOLDEN = Path(__file__).parent / "golden"


class _Pipe:
    device = "cuda"

    def __init__(self):
        self.calls = []

    def riffuse(self, inputs, init_image, mask_image=None):
        self.calls.append((inputs, init_image.size, None if mask_image is None else mask_image.mode))
        return init_image.copy()


def _seed_dir(tmp_path):
    rgb = np.load(GOLDEN / "og_beat.npz")["rgb"]
    Image.fromarray(rgb, mode="RGB").save(tmp_path / "og_beat.png")
    Image.fromarray(np.full((512, 512), 255, np.uint8), mode="L").save(tmp_path / "mask_all.png")
    return tmp_path


def test_compute_request_contract(tmp_path, monkeypatch):
    from riffusion import server
    from riffusion.datatypes import InferenceInput
    from riffusion.util.audio_segment import AudioSegment

    made = []

    class _Conv:
        def __init__(self, params, device):
            made.append((params, device))

        def audio_from_spectrogram_image(self, image, apply_filters=True):
            t = np.arange(int(44100 * 5.11))
            pcm = (3000 * np.sin(2 * np.pi * 440 * t / 44100)).astype(np.int16)[:, None]
            return AudioSegment(pcm, 44100)

    monkeypatch.setattr(server, "SpectrogramImageConverter", _Conv)
    seed = _seed_dir(tmp_path)
    pipe = _Pipe()
    payload = {"alpha": 0.25, "num_inference_steps": 50, "seed_image_id": "og_beat", "mask_image_id": "mask_all",
               "start": {"prompt": "church bells on sunday", "seed": 42}, "end": {"prompt": "jazz with piano", "seed": 123}}
    resp = server.run_inference(payload, pipe, seed)
    out = json.loads(resp)
    assert set(out) == {"image", "audio", "duration_s"} and abs(out["duration_s"] - 5.11) < 0.01
    assert out["image"].startswith("data:image/jpeg;base64,")
    img = Image.open(io.BytesIO(base64.decodebytes(out["image"].split(",", 1)[1].encode())))
    assert img.size == (512, 512) and img.format == "JPEG"
    assert out["audio"].startswith("data:audio/wav;base64,") or out["audio"].startswith("data:audio/mpeg;base64,")
    wav = base64.decodebytes(out["audio"].split(",", 1)[1].encode())
    assert wav[:4] == b"RIFF"
    inputs, size, mask_mode = pipe.calls[0]
    assert isinstance(inputs, InferenceInput) and inputs.alpha == 0.25 and inputs.start.denoising == 0.75
    assert size == (512, 512) and mask_mode == "RGB"                      # both images go through .convert("RGB") (:137,145)
    p, dev = made[0]
    assert (p.min_frequency, p.max_frequency, dev) == (0, 10000, "cuda")
    # error conventions: (message, 400)
    assert server.run_inference(dict(payload, seed_image_id="nope"), pipe, seed) == ("Invalid seed image: nope", 400)
    assert server.run_inference(dict(payload, mask_image_id="nope"), pipe, seed) == ("Invalid mask image: nope", 400)
    bad = server.run_inference({"alpha": 0.5, "start": {"prompt": "x"}}, pipe, seed)
    assert isinstance(bad, tuple) and bad[1] == 400
