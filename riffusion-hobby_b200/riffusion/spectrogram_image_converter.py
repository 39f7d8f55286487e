"""SpectrogramImageConverter — drop-in for riffusion/spectrogram_image_converter.py.

Thin composition of the image quantisation (riffusion.util.image_util) and SpectrogramConverter;
the audio arithmetic is in the CUDA kernels behind SpectrogramConverter.
"""
from __future__ import annotations

import numpy as np
from PIL import Image

from riffusion.spectrogram_converter import SpectrogramConverter
from riffusion.spectrogram_params import SpectrogramParams
from riffusion.util import image_util


class SpectrogramImageConverter:
    """Convert between spectrogram images and audio segments."""

    def __init__(self, params: SpectrogramParams, device: str = "cuda"):
        self.p = params
        self.device = device
        self.converter = SpectrogramConverter(params=params, device=device)

    def spectrogram_image_from_audio(self, segment) -> Image.Image:
        """AudioSegment -> spectrogram image with the conversion parameters and MAX_VALUE in its
        EXIF (spectrogram_image_converter.py:23-63)."""
        assert int(segment.frame_rate) == self.p.sample_rate, "Sample rate mismatch"

        if self.p.stereo:
            if segment.channels == 1:
                print("WARNING: Mono audio but stereo=True, cloning channel")
                segment = segment.set_channels(2)
            elif segment.channels > 2:
                print("WARNING: Multi channel audio, reducing to stereo")
                segment = segment.set_channels(2)
        elif segment.channels > 1:
            print("WARNING: Stereo audio but stereo=False, setting to mono")
            segment = segment.set_channels(1)

        spectrogram = self.converter.spectrogram_from_audio(segment)
        image = image_util.image_from_spectrogram(spectrogram, power=self.p.power_for_image)

        tags = self.p.to_exif()
        tags[SpectrogramParams.ExifTags.MAX_VALUE.value] = float(np.max(spectrogram))
        image.getexif().update(tags.items())
        return image

    def audio_from_spectrogram_image(self, image: Image.Image, apply_filters: bool = True, max_value: float = 30e6):
        """Spectrogram image -> AudioSegment (spectrogram_image_converter.py:65-91)."""
        spectrogram = image_util.spectrogram_from_image(
            image, max_value=max_value, power=self.p.power_for_image, stereo=self.p.stereo)
        return self.converter.audio_from_spectrogram(spectrogram, apply_filters=apply_filters)
