"""Programmatic dependent launch (csrc/rf_common.h) must not change a single bit: the same small UNet evaluation — eager
and as a replayed CUDA graph — is run in three processes with RF_PDL = 0 (attribute never set), 1 (the default: launches
that cannot fill the GPU) and 2 (every instrumented launch), and the outputs are compared byte for byte.  The mode is read
once per process, hence the subprocesses.  A dependent kernel that touched global memory before its griddepcontrol.wait
would show up here as a mismatch (or as garbage) in mode 1 / 2."""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent

_SCRIPT = r"""
import hashlib, sys
sys.path.insert(0, r"{root}")
sys.path.insert(0, r"{root}/riffusion-hobby_b200")
import torch
from oracle import unet_oracle as uo
from riffusion.graphed import GraphedUNet
from riffusion.unet_b200 import UNetB200

cfg = dict(block_out_channels=(64, 128, 128, 128), heads=4, cross_attention_dim=64)
oracle = uo.init_weights_(uo.UNet2DConditionOracle(**cfg), seed=0).eval()
g = torch.Generator().manual_seed(1)
lat = torch.randn(1, 4, 32, 32, generator=g).half().cuda()
ctx = torch.randn(2, 77, 64, generator=g).half().cuda()
with torch.no_grad():
    unet = UNetB200(oracle.state_dict(), device="cuda:0", block_out_channels=cfg["block_out_channels"], heads=4)
    x = torch.cat([lat, lat])
    outs = [unet(x, 741, encoder_hidden_states=ctx).sample for _ in range(3)]
    gr = GraphedUNet(unet, lat.shape, ctx)
    outs += [gr(lat, 741).clone() for _ in range(3)]
torch.cuda.synchronize()
assert all(torch.isfinite(o.float()).all() for o in outs)
assert all(torch.equal(outs[0], o) for o in outs[1:3]), "eager evaluations differ from run to run"
assert all(torch.equal(outs[3], o) for o in outs[4:]), "graph replays differ from run to run"
print("DIGEST", hashlib.sha256(outs[0].cpu().numpy().tobytes()).hexdigest(), hashlib.sha256(outs[3].cpu().numpy().tobytes()).hexdigest())
"""


def _run(mode: str) -> tuple[str, str]:
    env = dict(os.environ, RF_PDL=mode)
    r = subprocess.run([sys.executable, "-c", _SCRIPT.format(root=str(ROOT))], capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DIGEST")][-1].split()
    return line[1], line[2]


def test_pdl_modes_are_bit_identical(native_lib):
    ref = _run("0")
    for mode in ("1", "2"):
        assert _run(mode) == ref, f"RF_PDL={mode} changed the result"
