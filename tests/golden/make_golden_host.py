"""Golden vectors for the pure-host helpers, produced by IMPORTING the reference's own modules (they need only numpy /
torch / PIL, unlike the reference's pipeline module):

    python tests/golden/make_golden_host.py          # needs /root/reference; writes tests/golden/host_vectors.npz

  slerp_*      riffusion/util/torch_util.py:21-48 on seeded fp16 and fp32 tensors (numpy arithmetic in the tensors' dtype),
               including the nearly-parallel lerp branch
  prep_*       riffusion_pipeline.preprocess_image / preprocess_mask cannot be imported (diffusers), so they are covered by
               tests/test_pipeline_cpu.py against hand-computed values instead
Only numbers are stored — no reference source.
"""
from __future__ import annotations

import importlib.util
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _load(path: Path, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main() -> None:
    tu = _load(REF / "riffusion" / "util" / "torch_util.py", "ref_torch_util")
    out = {}
    torch.manual_seed(1234)
    a32, b32 = torch.randn(1, 4, 16, 16), torch.randn(1, 4, 16, 16)
    cases = {"f16": (a32.half(), b32.half()), "f32": (a32, b32),
             "par16": (a32.half(), (a32 * 1.0002 + 1e-3).half())}       # |cos| > 0.9995 -> lerp branch
    ts = np.array([0.0, 0.25, 0.5, 0.9, 1.0])
    out["ts"] = ts
    for name, (a, b) in cases.items():
        out[f"slerp_{name}_a"] = a.numpy()
        out[f"slerp_{name}_b"] = b.numpy()
        out[f"slerp_{name}_out"] = np.stack([tu.slerp(float(t), a, b).numpy() for t in ts])
    np.savez_compressed(OUT / "host_vectors.npz", **out)
    print("wrote", OUT / "host_vectors.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
