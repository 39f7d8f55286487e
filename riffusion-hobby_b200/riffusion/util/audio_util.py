"""Audio helpers (reference: riffusion/util/audio_util.py).

pydub is used when it is installed; otherwise the numpy AudioSegment stand-in from
`audio_segment.py` is (pydub/ffmpeg are absent from the B200 image).
"""
from __future__ import annotations

import io
import typing as T

import numpy as np
from scipy.io import wavfile

try:  # pragma: no cover - depends on the environment
    import pydub
    import pydub.effects

    AudioSegment = pydub.AudioSegment
    _normalize = pydub.effects.normalize
    HAVE_PYDUB = True
except ImportError:  # pragma: no cover
    from riffusion.util.audio_segment import AudioSegment, normalize as _normalize

    pydub = None
    HAVE_PYDUB = False


def int16_from_waveform(samples: np.ndarray, normalize: bool = False) -> np.ndarray:
    """(channels, samples) float -> (samples, channels) int16, audio_util.py:22-28:
    peak-normalise over all channels jointly to int16 max (in place, like the reference) and
    truncate toward zero."""
    if normalize:
        samples *= np.iinfo(np.int16).max / np.max(np.abs(samples))
    return samples.transpose(1, 0).astype(np.int16)


def audio_from_waveform(samples: np.ndarray, sample_rate: int, normalize: bool = False):
    """(channels, samples) float waveform -> AudioSegment via an in-memory WAV."""
    pcm = int16_from_waveform(samples, normalize=normalize)
    wav_bytes = io.BytesIO()
    wavfile.write(wav_bytes, sample_rate, pcm)
    wav_bytes.seek(0)
    return AudioSegment.from_wav(wav_bytes)


def apply_filters(segment, compression: bool = False):
    """Post-process a segment to about -12 dBFS with 0.1 dB peak headroom (audio_util.py:39-72)."""
    if compression:
        if not HAVE_PYDUB:
            raise NotImplementedError("compress_dynamic_range needs pydub (not installed)")
        segment = _normalize(segment, headroom=0.1)
        segment = segment.apply_gain(-10 - segment.dBFS)
        segment = pydub.effects.compress_dynamic_range(
            segment, threshold=-20.0, ratio=4.0, attack=5.0, release=50.0)
    desired_db = -12
    segment = segment.apply_gain(desired_db - segment.dBFS)
    return _normalize(segment, headroom=0.1)


def stitch_segments(segments: T.Sequence, crossfade_s: float):
    """Concatenate segments with a crossfade (audio_util.py:75-85)."""
    crossfade_ms = int(crossfade_s * 1000)
    combined = segments[0]
    for segment in segments[1:]:
        combined = combined.append(segment, crossfade=crossfade_ms)
    return combined


def overlay_segments(segments: T.Sequence):
    """Mix segments on top of each other (audio_util.py:88-99)."""
    assert len(segments) > 0
    output = None
    for segment in segments:
        output = segment if output is None else output.overlay(segment)
    return output
