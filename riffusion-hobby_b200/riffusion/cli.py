"""Command line tools — same commands and flags as the reference's argh CLI (riffusion/cli.py:21-278):

    python -m riffusion.cli audio-to-image --audio clip.wav --image clip.png
    python -m riffusion.cli image-to-audio --image clip.png --audio clip.wav
    python -m riffusion.cli print-exif --image clip.png
    python -m riffusion.cli sample-clips / audio-to-images-batch / sample-clips-batch ...

`argh` is not installed in the B200 image, so the sub-commands are declared once as keyword-only functions (as
in the reference) and exposed through a small argparse front end that derives `--flag-name` options from the
signatures.  Audio I/O uses pydub when present, else the WAV-only AudioSegment stand-in.
"""
from __future__ import annotations

import argparse
import inspect
import random
import sys
import typing as T
from multiprocessing.pool import ThreadPool
from pathlib import Path

import numpy as np
from PIL import Image

from riffusion.spectrogram_image_converter import SpectrogramImageConverter
from riffusion.spectrogram_params import SpectrogramParams
from riffusion.util import image_util
from riffusion.util.audio_util import AudioSegment


def audio_to_image(*, audio: str, image: str, step_size_ms: int = 10, num_frequencies: int = 512,
                   min_frequency: int = 0, max_frequency: int = 10000, window_duration_ms: int = 100,
                   padded_duration_ms: int = 400, power_for_image: float = 0.25, stereo: bool = False,
                   device: str = "cuda"):
    """Compute a spectrogram image from a waveform."""
    segment = AudioSegment.from_file(audio)
    params = SpectrogramParams(
        sample_rate=segment.frame_rate, stereo=stereo, window_duration_ms=window_duration_ms,
        padded_duration_ms=padded_duration_ms, step_size_ms=step_size_ms, min_frequency=min_frequency,
        max_frequency=max_frequency, num_frequencies=num_frequencies, power_for_image=power_for_image)
    converter = SpectrogramImageConverter(params=params, device=device)
    pil_image = converter.spectrogram_image_from_audio(segment)
    pil_image.save(image, exif=pil_image.getexif(), format="PNG")
    print(f"Wrote {image}")


def print_exif(*, image: str) -> None:
    """Print the params of a spectrogram image as saved in the exif data."""
    exif = image_util.exif_from_image(Image.open(image))
    for name, value in exif.items():
        print(f"{name:<20} = {value:>15}")


def image_to_audio(*, image: str, audio: str, device: str = "cuda"):
    """Reconstruct an audio clip from a spectrogram image."""
    pil_image = Image.open(image)
    img_exif = pil_image.getexif()
    assert img_exif is not None
    try:
        params = SpectrogramParams.from_exif(exif=img_exif)
    except KeyError:
        print("WARNING: Could not find spectrogram parameters in exif data. Using defaults.")
        params = SpectrogramParams()
    converter = SpectrogramImageConverter(params=params, device=device)
    segment = converter.audio_from_spectrogram_image(pil_image)
    extension = Path(audio).suffix[1:]
    segment.export(audio, format=extension)
    print(f"Wrote {audio} ({segment.duration_seconds:.2f} seconds)")


def sample_clips(*, audio: str, output_dir: str, num_clips: int = 1, duration_ms: int = 5120, mono: bool = False,
                 extension: str = "wav", seed: int = -1):
    """Slice an audio file into clips of the given duration."""
    if seed >= 0:
        np.random.seed(seed)
    segment = AudioSegment.from_file(audio)
    if mono:
        segment = segment.set_channels(1)
    out_dir = Path(output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    total_ms = int(segment.duration_seconds * 1000)
    for i in range(num_clips):
        start = np.random.randint(0, total_ms - duration_ms)
        clip = segment[start: start + duration_ms]
        path = out_dir / f"clip_{i}_start_{start}_ms_duration_{duration_ms}_ms.{extension}"
        clip.export(path, format=extension)
        print(f"Wrote {path}")


def audio_to_images_batch(*, audio_dir: str, output_dir: str, image_extension: str = "jpg", step_size_ms: int = 10,
                          num_frequencies: int = 512, min_frequency: int = 0, max_frequency: int = 10000,
                          power_for_image: float = 0.25, mono: bool = False, sample_rate: int = 44100,
                          device: str = "cuda", num_threads: T.Optional[int] = None, limit: int = -1):
    """Process audio clips into spectrograms in batch, multi-threaded (one converter shared by all threads)."""
    audio_paths = sorted(p for p in Path(audio_dir).glob("*") if p.is_file())
    if limit > 0:
        audio_paths = audio_paths[:limit]
    out_dir = Path(output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    params = SpectrogramParams(step_size_ms=step_size_ms, num_frequencies=num_frequencies, min_frequency=min_frequency,
                               max_frequency=max_frequency, stereo=not mono, power_for_image=power_for_image,
                               sample_rate=sample_rate)
    converter = SpectrogramImageConverter(params=params, device=device)

    def process_one(audio_path: Path) -> None:
        try:
            segment = AudioSegment.from_file(str(audio_path))
        except Exception:  # noqa: BLE001 - unreadable files are skipped like in the reference (cli.py:176-179)
            return
        if mono and segment.channels != 1:
            segment = segment.set_channels(1)
        elif not mono and segment.channels != 2:
            segment = segment.set_channels(2)
        if segment.frame_rate != params.sample_rate:
            segment = segment.set_frame_rate(params.sample_rate)
        image = converter.spectrogram_image_from_audio(segment)
        image_path = out_dir / f"{audio_path.stem}.{image_extension}"
        fmt = {"jpg": "JPEG", "jpeg": "JPEG", "png": "PNG"}[image_extension]
        image.save(image_path, exif=image.getexif(), format=fmt)

    with ThreadPool(processes=num_threads) as pool:
        for _ in pool.imap_unordered(process_one, audio_paths):
            pass


def sample_clips_batch(*, audio_dir: str, output_dir: str, num_clips_per_file: int = 1, duration_ms: int = 5120,
                       mono: bool = False, extension: str = "mp3", num_threads: T.Optional[int] = None, glob: str = "*",
                       limit: int = -1, seed: int = -1):
    """Sample short clips from a directory of audio files, multi-threaded."""
    audio_paths = sorted(p for p in Path(audio_dir).glob(glob) if p.is_file())
    if limit > 0:
        audio_paths = audio_paths[:limit]
    if seed >= 0:
        random.seed(seed)
    out_dir = Path(output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)

    def process_one(audio_path: Path) -> None:
        try:
            segment = AudioSegment.from_file(str(audio_path))
        except Exception:  # noqa: BLE001
            return
        if mono:
            segment = segment.set_channels(1)
        total_ms = int(segment.duration_seconds * 1000)
        for i in range(num_clips_per_file):
            start = np.random.randint(0, max(total_ms - duration_ms, 1))
            clip = segment[start: start + duration_ms]
            name = f"{audio_path.stem}_{i}_start_{start}_ms_dur_{duration_ms}_ms.{extension}"
            clip.export(out_dir / name, format=extension)

    with ThreadPool(processes=num_threads) as pool:
        for _ in pool.imap_unordered(process_one, audio_paths):
            pass


COMMANDS = [audio_to_image, image_to_audio, sample_clips, print_exif, audio_to_images_batch, sample_clips_batch]


def _str2bool(v: str) -> bool:
    if v.lower() in ("1", "true", "yes", "y"):
        return True
    if v.lower() in ("0", "false", "no", "n"):
        return False
    raise argparse.ArgumentTypeError(f"expected a boolean, got {v!r}")


def build_parser() -> argparse.ArgumentParser:
    """argh-style front end: one sub-command per function (underscores -> dashes), one --flag per keyword-only arg."""
    parser = argparse.ArgumentParser(prog="riffusion.cli", description=__doc__)
    sub = parser.add_subparsers(dest="command", required=True)
    for fn in COMMANDS:
        sp = sub.add_parser(fn.__name__.replace("_", "-"), help=(fn.__doc__ or "").strip())
        sp.set_defaults(_fn=fn)
        for name, prm in inspect.signature(fn).parameters.items():
            flag = "--" + name.replace("_", "-")
            if prm.default is inspect.Parameter.empty:
                sp.add_argument(flag, dest=name, required=True)
            elif isinstance(prm.default, bool):
                sp.add_argument(flag, dest=name, nargs="?", const=True, default=prm.default, type=_str2bool)
            elif prm.default is None:
                sp.add_argument(flag, dest=name, default=None, type=int)
            else:
                sp.add_argument(flag, dest=name, default=prm.default, type=type(prm.default))
    return parser


def main(argv: T.Optional[T.Sequence[str]] = None) -> None:
    args = vars(build_parser().parse_args(argv))
    fn = args.pop("_fn")
    args.pop("command")
    fn(**args)


if __name__ == "__main__":
    main(sys.argv[1:])
