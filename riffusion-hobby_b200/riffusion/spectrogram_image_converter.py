"""Images <-> audio: the composition layer over SpectrogramConverter.

API-compatible with the reference module of the same name (class `SpectrogramImageConverter`, attributes `p`, `device`,
`converter`, methods `spectrogram_image_from_audio` / `audio_from_spectrogram_image`; reference file
riffusion/spectrogram_image_converter.py:10-91).  Nothing numerical happens here: quantisation lives in
riffusion.util.image_util (host) / rf_image_to_mel, rf_mel_to_image (device), the audio arithmetic in the CUDA kernels
behind SpectrogramConverter.
"""
from __future__ import annotations

import numpy as np
from PIL import Image

from riffusion.spectrogram_converter import SpectrogramConverter
from riffusion.spectrogram_params import SpectrogramParams
from riffusion.util import image_util

_MAX_VALUE_TAG = SpectrogramParams.ExifTags.MAX_VALUE.value


def _conform_channels(segment, want_stereo: bool):
    """Channel policy of the reference (:37-47): mono is cloned to stereo, anything wider than stereo is folded down,
    stereo is mixed down when the parameters say mono.  The warnings are the reference's."""
    have = segment.channels
    if want_stereo and have != 2:
        print("WARNING: Mono audio but stereo=True, cloning channel" if have == 1
              else "WARNING: Multi channel audio, reducing to stereo")
        return segment.set_channels(2)
    if not want_stereo and have > 1:
        print("WARNING: Stereo audio but stereo=False, setting to mono")
        return segment.set_channels(1)
    return segment


class SpectrogramImageConverter:
    """Spectrogram image <-> audio segment, for one set of SpectrogramParams on one device."""

    def __init__(self, params: SpectrogramParams, device: str = "cuda"):
        self.p = params
        self.device = device
        self.converter = SpectrogramConverter(params=params, device=device)

    def spectrogram_image_from_audio(self, segment) -> Image.Image:
        """Audio segment -> PIL image whose EXIF carries the conversion parameters plus MAX_VALUE, the largest mel
        amplitude, which a reader needs to undo the 8-bit normalisation (:23-63)."""
        if int(segment.frame_rate) != self.p.sample_rate:
            raise AssertionError("Sample rate mismatch")
        amplitudes = self.converter.spectrogram_from_audio(_conform_channels(segment, self.p.stereo))
        picture = image_util.image_from_spectrogram(amplitudes, power=self.p.power_for_image)
        picture.getexif().update({**self.p.to_exif(), _MAX_VALUE_TAG: float(np.max(amplitudes))}.items())
        return picture

    def audio_from_spectrogram_image(self, image: Image.Image, apply_filters: bool = True, max_value: float = 30e6):
        """PIL spectrogram image -> audio segment (:65-91).  `max_value` rescales the 8-bit amplitudes; the output is
        peak-normalised afterwards, so its exact value does not matter."""
        amplitudes = image_util.spectrogram_from_image(image, max_value=max_value, power=self.p.power_for_image,
                                                       stereo=self.p.stereo)
        return self.converter.audio_from_spectrogram(amplitudes, apply_filters=apply_filters)
