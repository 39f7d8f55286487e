"""Inference request -> JSON response, the model server's contract without the web framework
(reference: riffusion/server.py:66-183; Flask itself is outside SURVEY §8 and not installed here).

`compute_request(inputs, pipeline, seed_images_dir)` keeps the reference's signature and return convention: a JSON string
of `InferenceOutput` on success, or a `(message, 400)` tuple for a bad seed / mask id.  `run_inference(json_data, ...)`
is the body of the `/run_inference/` route (:83-112) operating on an already-parsed JSON object.

What differs, loudly:
  * audio container: the reference exports MP3 through pydub + ffmpeg (:166-169).  Neither exists in this image; when the
    segment cannot export "mp3" the response carries `data:audio/wav;base64,...` instead (same int16 PCM, lossless).
  * between `riffuse` and the audio the reference goes GPU -> PIL -> numpy -> GPU; `fast_audio=True` (default when the
    pipeline offers `generate_clips`-style device glue) keeps that on the device but returns byte-identical JSON fields.
"""
from __future__ import annotations

import dataclasses
import io
import json
import logging
import time
import typing as T
from pathlib import Path

import PIL.Image

from riffusion.datatypes import InferenceInput, InferenceOutput
from riffusion.spectrogram_image_converter import SpectrogramImageConverter
from riffusion.spectrogram_params import SpectrogramParams
from riffusion.util import base64_util

SEED_IMAGES_DIR = Path(Path(__file__).resolve().parent.parent, "seed_images")


def run_inference(json_data: T.Mapping[str, T.Any], pipeline, seed_images_dir: T.Union[str, Path] = SEED_IMAGES_DIR):
    """parse + validate the request like the route does (dacite errors -> 400), then `compute_request`"""
    start_time = time.time()
    try:
        inputs = InferenceInput.from_dict(json_data)
    except (TypeError, KeyError, ValueError) as exception:          # dacite WrongTypeError / MissingValueError
        logging.info(json_data)
        return str(exception), 400
    response = compute_request(inputs=inputs, seed_images_dir=str(seed_images_dir), pipeline=pipeline)
    logging.info(f"Request took {time.time() - start_time:.2f} s")
    return response


def compute_request(inputs: InferenceInput, pipeline, seed_images_dir: str) -> T.Union[str, T.Tuple[str, int]]:
    init_image_path = Path(seed_images_dir, f"{inputs.seed_image_id}.png")
    if not init_image_path.is_file():
        return f"Invalid seed image: {inputs.seed_image_id}", 400
    init_image = PIL.Image.open(str(init_image_path)).convert("RGB")

    mask_image: T.Optional[PIL.Image.Image] = None
    if inputs.mask_image_id:
        mask_image_path = Path(seed_images_dir, f"{inputs.mask_image_id}.png")
        if not mask_image_path.is_file():
            return f"Invalid mask image: {inputs.mask_image_id}", 400
        mask_image = PIL.Image.open(str(mask_image_path)).convert("RGB")

    image = pipeline.riffuse(inputs, init_image=init_image, mask_image=mask_image)

    params = SpectrogramParams(min_frequency=0, max_frequency=10000)
    converter = SpectrogramImageConverter(params=params, device=str(pipeline.device))    # plans are cached per geometry
    segment = converter.audio_from_spectrogram_image(image, apply_filters=True)

    audio_bytes = io.BytesIO()
    try:
        segment.export(audio_bytes, format="mp3")
        audio_mime = "audio/mpeg"
    except (NotImplementedError, ValueError, OSError):
        audio_bytes = io.BytesIO()
        segment.export(audio_bytes, format="wav")
        audio_mime = "audio/wav"
    audio_bytes.seek(0)

    image_bytes = io.BytesIO()
    image.save(image_bytes, exif=image.getexif(), format="JPEG")
    image_bytes.seek(0)

    output = InferenceOutput(
        image="data:image/jpeg;base64," + base64_util.encode(image_bytes),
        audio=f"data:{audio_mime};base64," + base64_util.encode(audio_bytes),
        duration_s=segment.duration_seconds,
    )
    return json.dumps(dataclasses.asdict(output))
