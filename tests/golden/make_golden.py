"""Generate the golden fixtures under tests/golden/ from the reference checkout.

Run once in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

Outputs
  tired_traveler_clip2.npz  the reference's own known-answer pair for the forward path
      (test/test_data/tired_traveler/clips/clip_2_*.wav  ->  images/clip_2_*_stereo.png, incl. the
      EXIF tags written by spectrogram_image_converter.py:58-61, and the mono PNG)
  og_beat.npz               seed_images/og_beat.png after PIL's .convert("RGB") (BASELINE config 1/2 input)
  torchaudio_vectors.npz    small seeded input/output vectors produced by the installed torchaudio
      transforms built with the reference's arguments (oracle/torchaudio_ref.py): inverse mel and
      Griffin-Lim with recorded initial angles, STFT+mel of a short waveform.
Only data is copied (audio samples, pixels, EXIF numbers) — no reference source.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
from PIL import Image
from scipy.io import wavfile

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def exif_dict(img: Image.Image) -> dict:
    return {int(k): v for k, v in img.getexif().items()}


def main() -> None:
    tt = REF / "test" / "test_data" / "tired_traveler"
    rate, wav = wavfile.read(tt / "clips" / "clip_2_start_103694_ms_duration_5678_ms.wav")
    stereo = Image.open(tt / "images" / "clip_2_start_103694_ms_duration_5678_ms_stereo.png")
    mono = Image.open(tt / "images" / "clip_2_start_103694_ms_duration_5678_ms.png")
    ex_s, ex_m = exif_dict(stereo), exif_dict(mono)
    keys = sorted(ex_s)
    np.savez_compressed(
        OUT / "tired_traveler_clip2.npz",
        wav=wav, rate=np.int64(rate),
        stereo_png=np.array(stereo.convert("RGB")), mono_png=np.array(mono.convert("RGB")),
        stereo_mode=np.array(stereo.mode), mono_mode=np.array(mono.mode),
        exif_keys=np.array(keys, np.int64), exif_stereo=np.array([float(ex_s[k]) for k in keys]),
        exif_mono=np.array([float(ex_m.get(k, np.nan)) for k in keys]),
    )
    og = Image.open(REF / "seed_images" / "og_beat.png")
    np.savez_compressed(OUT / "og_beat.npz", rgb=np.array(og.convert("RGB")), mode=np.array(og.mode),
                        n_exif=np.int64(len(og.getexif())))

    from oracle.torchaudio_ref import TorchaudioConverter, griffinlim_with_angles

    torch.manual_seed(1234)
    T_ = 24
    conv = TorchaudioConverter(n_iter=4)
    mel = (torch.rand(1, 512, T_) ** 4) * 3e7
    lin = conv.inverse_mel_scaler(mel)
    angles = torch.rand(1, 8821, T_, dtype=torch.complex64)
    wave = griffinlim_with_angles(conv.inverse_spectrogram_func, lin, angles)
    x = torch.randn(1, 12000) * 3000
    mel_fwd = conv.mel_amplitudes_from_waveform(x)
    live = (conv.mel_scaler.fb != 0).any(dim=1).numpy()
    np.savez_compressed(
        OUT / "torchaudio_vectors.npz",
        mel=mel.numpy(), lin_live=lin.numpy()[:, live], live=live,
        angles_live=angles.numpy()[:, live], wave=wave.numpy(), n_iter=np.int64(4),
        x=x.numpy(), mel_fwd=mel_fwd.numpy(),
        torch_version=np.array(torch.__version__),
    )
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
