"""Parameters of the audio <-> spectrogram <-> image conversion.

API-compatible with the reference dataclass (riffusion/spectrogram_params.py:8-115): same
field names and defaults, the same custom EXIF tag ids (11000-11080), the same derived
`n_fft` / `win_length` / `hop_length`, and `to_exif` / `from_exif`.  The object is frozen
and hashable (the reference's Streamlit layer caches on it, streamlit/util.py:188-192).
"""
from __future__ import annotations

import enum
import typing as T
from dataclasses import dataclass, fields


def _ms_to_samples(duration_ms: int, sample_rate: int) -> int:
    # int(ms / 1000.0 * sr): float product truncated, exactly as the reference derives sizes
    return int(duration_ms / 1000.0 * sample_rate)


@dataclass(frozen=True)
class SpectrogramParams:
    # channel layout
    stereo: bool = False

    # STFT geometry
    sample_rate: int = 44100
    step_size_ms: int = 10
    window_duration_ms: int = 100
    padded_duration_ms: int = 400

    # mel scale
    num_frequencies: int = 512
    min_frequency: int = 0
    max_frequency: int = 10000
    mel_scale_norm: T.Optional[str] = None
    mel_scale_type: str = "htk"
    max_mel_iters: int = 200

    # phase reconstruction
    num_griffin_lim_iters: int = 32

    # image curve
    power_for_image: float = 0.25

    class ExifTags(enum.Enum):
        """Custom EXIF tag ids carried by spectrogram images."""

        SAMPLE_RATE = 11000
        STEREO = 11005
        STEP_SIZE_MS = 11010
        WINDOW_DURATION_MS = 11020
        PADDED_DURATION_MS = 11030

        NUM_FREQUENCIES = 11040
        MIN_FREQUENCY = 11050
        MAX_FREQUENCY = 11060

        POWER_FOR_IMAGE = 11070
        MAX_VALUE = 11080

    @property
    def n_fft(self) -> int:
        """Samples per (zero-padded) STFT frame."""
        return _ms_to_samples(self.padded_duration_ms, self.sample_rate)

    @property
    def win_length(self) -> int:
        """Samples under the analysis window."""
        return _ms_to_samples(self.window_duration_ms, self.sample_rate)

    @property
    def hop_length(self) -> int:
        """Samples between consecutive frames."""
        return _ms_to_samples(self.step_size_ms, self.sample_rate)

    # (field name, tag, caster applied when writing / reading)
    _EXIF_FIELDS: T.ClassVar[T.Tuple[T.Tuple[str, str, T.Callable], ...]] = (
        ("sample_rate", "SAMPLE_RATE", lambda v: v),
        ("stereo", "STEREO", bool),
        ("step_size_ms", "STEP_SIZE_MS", lambda v: v),
        ("window_duration_ms", "WINDOW_DURATION_MS", lambda v: v),
        ("padded_duration_ms", "PADDED_DURATION_MS", lambda v: v),
        ("num_frequencies", "NUM_FREQUENCIES", lambda v: v),
        ("min_frequency", "MIN_FREQUENCY", lambda v: v),
        ("max_frequency", "MAX_FREQUENCY", lambda v: v),
        ("power_for_image", "POWER_FOR_IMAGE", lambda v: v),
    )

    def to_exif(self) -> T.Dict[int, T.Any]:
        """EXIF dictionary {tag id: value} for these parameters (MAX_VALUE is added by the
        image converter)."""
        out: T.Dict[int, T.Any] = {}
        for name, tag, _ in self._EXIF_FIELDS:
            value = getattr(self, name)
            if name == "power_for_image":
                value = float(value)
            out[self.ExifTags[tag].value] = value
        return out

    @classmethod
    def from_exif(cls, exif: T.Mapping[int, T.Any]) -> "SpectrogramParams":
        """Inverse of `to_exif`; raises KeyError when a tag is missing (the CLI relies on
        that to fall back to defaults, cli.py:80-87)."""
        kwargs = {name: cast(exif[cls.ExifTags[tag].value]) for name, tag, cast in cls._EXIF_FIELDS}
        return cls(**kwargs)


assert {f.name for f in fields(SpectrogramParams)} >= {n for n, _, _ in SpectrogramParams._EXIF_FIELDS}
