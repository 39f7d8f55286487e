"""Golden vectors for row a-1 / a-6 (image <-> spectrogram), produced by IMPORTING the reference's own
riffusion/util/image_util.py (it needs only numpy + PIL + the reference's dataclass module):

    python tests/golden/make_golden_image.py        # needs /root/reference; writes tests/golden/image_vectors.npz

  sfi_*   image_util.spectrogram_from_image (:59-110) on a 96x80 crop of seed_images/og_beat.png (the RGB array the
          test feeds is stored too), mono and stereo, and on an "L"-mode version of the crop (P/L -> RGB branch :81-82)
  ifs_*   image_util.image_from_spectrogram (:13-56) on a seeded 2-channel and 1-channel amplitude array
Only numbers are stored — no reference source.  Run in a fresh interpreter: the reference package is named `riffusion`
like ours, so only /root/reference may be on sys.path.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference")
from PIL import Image  # noqa: E402

from riffusion.util import image_util as ref  # noqa: E402  (the reference's)

assert ref.__file__.startswith("/root/reference"), ref.__file__
OUT = Path(__file__).resolve().parent


def main() -> None:
    out = {}
    og = Image.open("/root/reference/seed_images/og_beat.png")
    rgb = np.array(og.convert("RGB"))[200:296, 100:180].copy()          # (96, 80, 3)
    out["sfi_rgb"] = rgb
    pil = Image.fromarray(rgb, mode="RGB")
    out["sfi_mono"] = ref.spectrogram_from_image(pil, power=0.25, stereo=False, max_value=30e6)
    out["sfi_stereo"] = ref.spectrogram_from_image(pil, power=0.25, stereo=True, max_value=30e6)
    out["sfi_mono_p05"] = ref.spectrogram_from_image(pil, power=0.5, stereo=False, max_value=1234.5)
    gray = pil.convert("L")
    out["sfi_gray_u8"] = np.array(gray)
    out["sfi_gray_mono"] = ref.spectrogram_from_image(gray, power=0.25, stereo=False, max_value=30e6)
    rng = np.random.default_rng(7)
    spec2 = (rng.random((2, 48, 40), dtype=np.float32) ** 4 * np.float32(4.6e7)).astype(np.float32)
    spec1 = spec2[:1].copy()
    out["ifs_spec2"] = spec2
    out["ifs_img2"] = np.array(ref.image_from_spectrogram(spec2, power=0.25))
    out["ifs_img1"] = np.array(ref.image_from_spectrogram(spec1, power=0.25))
    np.savez_compressed(OUT / "image_vectors.npz", **out)
    print("wrote", OUT / "image_vectors.npz", {k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
