"""GPU parity of the path that bench.py times (VERDICT r1 "next round" items 1-2): full-size SD-1.5 UNet at the benchmarked
CFG batch 64, CUDA-graph replay vs eager at that size, VAE decode at 64x64 latents, the fp16 image -> uint8 step,
`RiffusionPipeline.generate_clips` end to end (denoise -> VAE -> uint8 -> mel -> inverse mel + Griffin-Lim), and
`RiffusionPipeline.riffuse()` PIL -> PIL with a mask (BASELINE configs[2]).

Checkers (test infrastructure, never the product): oracle/unet_oracle.py + oracle/vae_oracle.py in fp32 (UNPINNED
restatement of diffusers 0.9: diffusers is not installable), oracle/unet_emul.py (the same modules with fp16 STORAGE
at the kernels' rounding points), oracle/audio_oracle.py and the installed torchaudio transforms.

Tolerance of whole-network outputs.  north_star: "latents within 1e-3 relative fp16".  That bar is below the noise of
fp16 storage for this network, for ANY implementation: the emulation (fp32 math, fp16 rounding at the kernels' storage
points) sits 1.4e-3 from the fp32 oracle (the "floor"), and a SECOND emulation that differs only below fp32 rounding
(contractions in float64, identical rounding points) sits 1.7-2.0e-3 from the first — rounding decisions decorrelate
through ~100 layers, so two correct fp16 evaluations are sqrt(2) x floor apart and neither can be within 1e-3 of the
other unless it is bit-identical.  Both numbers are measured and printed by the tests below.  The bars therefore are
    rel_l2(kernels, fp32 oracle)  <= 1.15 * floor + 1e-4      (nothing beyond what fp16 storage itself costs; measured:
                                                               the kernels sit AT the floor, 1.42e-3 vs 1.43e-3)
    rel_l2(kernels, emulation)    <= 1.25 * rel_l2(emulation', emulation) + 1e-4
                                                              (as close to one fp16 evaluation as another one is)
For multi-step loops the same comparisons are made against the fp32 oracle loop and the emulated loop(s).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _round_params(m):
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(p.half().float())
    return m


@pytest.fixture(scope="module")
def sd15(native_lib):
    """full-size SD-1.5 UNet: fp32 oracle module (fp16-representable random-init weights) + UNetB200 on the same weights"""
    from oracle import unet_oracle as uo
    from riffusion.unet_b200 import UNetB200

    oracle = _round_params(uo.init_weights_(uo.UNet2DConditionOracle(), seed=0)).cuda().eval()
    return oracle, UNetB200(oracle.state_dict(), device="cuda")


@pytest.fixture(scope="module")
def small_unet(native_lib):
    from oracle import unet_oracle as uo
    from riffusion.unet_b200 import UNetB200

    cfg = dict(block_out_channels=(64, 128, 128, 128), heads=4, cross_attention_dim=64)
    oracle = _round_params(uo.init_weights_(uo.UNet2DConditionOracle(**cfg), seed=21)).cuda().eval()
    return oracle, UNetB200(oracle.state_dict(), device="cuda", block_out_channels=cfg["block_out_channels"], heads=4)


@pytest.fixture(scope="module")
def vae_pair(native_lib):
    from oracle.unet_oracle import init_weights_
    from oracle.vae_oracle import AutoencoderKLOracle
    from riffusion.vae_b200 import VaeB200

    oracle = _round_params(init_weights_(AutoencoderKLOracle(), seed=5, std=0.03)).cuda().eval()
    return oracle, VaeB200(oracle.state_dict(), device="cuda")


@pytest.fixture(scope="module")
def conv(native_lib):
    from riffusion.spectrogram_converter import SpectrogramConverter
    from riffusion.spectrogram_params import SpectrogramParams

    return SpectrogramConverter(SpectrogramParams(), device="cuda")


def _check_vs_floor(got, ref32, emul, what, emul2=None, floor_factor=1.15):
    """got: kernels; ref32: fp32 oracle; emul: fp16-storage emulation; emul2: the second fp16-storage evaluation (float64
    contractions, same rounding points) or None"""
    e_k, e_f, e_o = rel_l2(got, emul), rel_l2(emul, ref32), rel_l2(got, ref32)
    e_2 = rel_l2(emul2, emul) if emul2 is not None else None
    print(f"{what}: kernels vs fp32 oracle {e_o:.3e} | fp16-storage floor (emulation vs fp32) {e_f:.3e} | kernels vs emulation "
          f"{e_k:.3e} | two fp16-storage evaluations apart " + (f"{e_2:.3e}" if e_2 is not None else "n/a"))
    assert torch.isfinite(got.float()).all()
    assert e_o <= floor_factor * e_f + 1e-4, f"{what}: kernels vs fp32 {e_o:.3e}, floor {e_f:.3e}"
    spread = e_2 if e_2 is not None else 2 ** 0.5 * floor_factor * e_f       # independent rounding noise adds in quadrature
    assert e_k <= 1.25 * spread + 1e-4, f"{what}: kernels vs emulation {e_k:.3e}, spread of fp16 evaluations {spread:.3e}"
    return e_k, e_f, e_o


def _emul_hi(fn, *a, **k):
    from oracle import unet_emul as ue

    ue.HI = True
    try:
        return fn(*a, **k)
    finally:
        ue.HI = False


# ----------------------------------------------------------------------------------------------- UNet
@torch.no_grad()
def test_unet_full_size_fp16_storage_floor(sd15):
    """B = 2 (one CFG pair): the three distances that define the tolerance, at two timesteps"""
    from oracle import unet_emul as ue

    oracle, ours = sd15
    for t in (741, 21):
        torch.manual_seed(t)
        x = torch.randn(2, 4, 64, 64, device="cuda").half()
        ctx = torch.randn(2, 77, 768, device="cuda").half()
        ref32 = oracle(x.float(), t, ctx.float())
        emul = ue.unet_forward(oracle, x, t, ctx)
        emul2 = _emul_hi(ue.unet_forward, oracle, x, t, ctx)
        got = ours(x, t, encoder_hidden_states=ctx).sample
        _check_vs_floor(got, ref32, emul, f"UNet SD-1.5 64x64 B=2 t={t}", emul2=emul2)


@torch.no_grad()
def test_unet_benchmarked_batch64_matches_oracle_and_graph(sd15):
    """bench.py's shape: 32 requests -> CFG batch 64 ([uncond x 32 | text x 32], per-request text embeddings), the
    tile / split-K choices of that batch.  Oracle evaluated on a subset of the 64 images (they are independent);
    then the CUDA-graph replay of the same evaluation must equal the eager launch sequence bit for bit."""
    from oracle import unet_emul as ue
    from riffusion.graphed import GraphedUNet

    oracle, ours = sd15
    torch.manual_seed(64)
    lat = torch.randn(32, 4, 64, 64, device="cuda").half()
    text = torch.randn(32, 77, 768, device="cuda").half()
    uncond = torch.randn(1, 77, 768, device="cuda").half()
    ctx = torch.cat([uncond.expand(32, -1, -1), text]).contiguous()
    x = torch.cat([lat, lat])
    t = 961
    got = ours(x, t, encoder_hidden_states=ctx).sample
    assert got.shape == (64, 4, 64, 64)
    idx = torch.tensor([0, 13, 31, 32, 45, 63], device="cuda")
    ref32 = torch.cat([oracle(x[i:i + 1].float(), t, ctx[i:i + 1].float()) for i in idx.tolist()])
    emul = torch.cat([ue.unet_forward(oracle, x[i:i + 1], t, ctx[i:i + 1]) for i in idx.tolist()])
    emul2 = torch.cat([_emul_hi(ue.unet_forward, oracle, x[i:i + 1], t, ctx[i:i + 1]) for i in idx.tolist()[:2]])
    print(f"two fp16-storage evaluations apart (2 images): {rel_l2(emul2, emul[:2]):.3e}")
    _check_vs_floor(got[idx], ref32, emul, "UNet SD-1.5 CFG batch 64 (subset of 6 images)")
    # images are independent, but a batch of 2 takes other tile shapes / split-K than a batch of 64: a different
    # accumulation order below fp32 rounding, i.e. one more "equally valid fp16 evaluation" — it sits as far from the
    # batch-64 result as two emulations sit from each other, not closer
    pair = ours(x[[0, 32]].contiguous(), t, encoder_hidden_states=ctx[[0, 32]].contiguous()).sample
    spread = rel_l2(emul2, emul[:2])
    e_pair = rel_l2(pair, got[[0, 32]])
    print(f"same images at batch 2 vs batch 64: {e_pair:.3e} (spread of two fp16-storage evaluations {spread:.3e})")
    assert e_pair <= 1.25 * spread + 1e-4
    assert torch.equal(ours(x[[0, 32]].contiguous(), t, encoder_hidden_states=ctx[[0, 32]].contiguous()).sample, pair)   # deterministic
    graphed = GraphedUNet(ours, lat.shape, ctx)
    g = graphed(lat, t).clone()
    assert torch.equal(g, got), "CUDA-graph replay differs from the eager evaluation at full size"
    g2 = graphed(lat, t)
    assert torch.equal(g2, got)


# ----------------------------------------------------------------------------------------------- VAE + uint8
@torch.no_grad()
def test_vae_decode_64x64_latents_and_uint8(vae_pair):
    """b-6 at the benchmarked size: (B,4,64,64) latents -> (B,3,512,512); then riffusion_pipeline.py:430-434
    (fp16 `image / 2 + 0.5`, clamp, float16 `* 255`, round) must be reproduced EXACTLY by rf_vae_image_to_u8."""
    from oracle import unet_emul as ue
    from oracle.vae_oracle import u8_from_image_fp16
    from riffusion import tc_ops as ops

    oracle, ours = vae_pair
    torch.manual_seed(66)
    z = (torch.randn(2, 4, 64, 64, device="cuda") * 0.18215 * 4).half()
    zs = ((1.0 / 0.18215) * z)                                         # fp16, as the reference rescales (:427)
    got = ours.decode(zs).sample
    assert got.shape == (2, 3, 512, 512) and got.dtype == torch.float16
    ref32 = torch.cat([oracle.decode(zs[i:i + 1].float()) for i in range(2)])
    emul = torch.cat([ue.vae_decode(oracle, zs[i:i + 1]) for i in range(2)])
    emul2 = _emul_hi(ue.vae_decode, oracle, zs[0:1])
    _check_vs_floor(got[0:1], ref32[0:1], emul[0:1], "VAE decode 64x64 latents (image 0)", emul2=emul2)
    _check_vs_floor(got, ref32, emul, "VAE decode 64x64 latents")
    # exact uint8 step, on the decoded image and on a synthetic image that covers the clamp and every rounding tie
    for img in (got, (torch.rand(1, 3, 64, 512, device="cuda") * 2.6 - 1.3).half(),
                torch.linspace(-1.2, 1.2, 3 * 256 * 256, device="cuda").reshape(1, 3, 256, 256).half()):
        u8 = ops.vae_image_to_u8(img)
        want = u8_from_image_fp16(img)
        assert u8.shape == want.shape and u8.dtype == torch.uint8
        assert np.array_equal(u8.cpu().numpy(), want), "rf_vae_image_to_u8 is not bit-exact"


# ----------------------------------------------------------------------------------------------- a-1 vs oracle / reference
def test_image_to_mel_matches_oracle_and_reference_vectors(native_lib):
    """row a-1: rf_image_to_mel against (i) vectors produced by importing the reference's own
    riffusion/util/image_util.py:59-110 (tests/golden/make_golden_image.py) and (ii) the oracle restatement on the full
    og_beat image; mono = R plane, stereo = G,B planes, flip-Y, x^(1/power) * max_value.  a-6 the other way round."""
    from pathlib import Path

    from oracle import audio_oracle as ao
    from riffusion.util import image_util

    g = np.load(Path(__file__).parent / "golden" / "image_vectors.npz")
    rgb = torch.from_numpy(g["sfi_rgb"]).cuda()
    for key, kw in (("sfi_mono", dict(power=0.25, stereo=False, max_value=30e6)),
                    ("sfi_stereo", dict(power=0.25, stereo=True, max_value=30e6)),
                    ("sfi_mono_p05", dict(power=0.5, stereo=False, max_value=1234.5))):
        got = image_util.spectrogram_from_image_device(rgb, **kw).cpu().numpy()
        want = g[key]
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.abs(got - want).max() <= 4e-6 * want.max(), key       # device powf vs numpy power
        zero = want == 0
        assert np.all(got[zero] == 0)
    gray = torch.from_numpy(np.repeat(g["sfi_gray_u8"][:, :, None], 3, axis=2).copy()).cuda()   # PIL L -> RGB replicates
    got = image_util.spectrogram_from_image_device(gray, power=0.25, stereo=False, max_value=30e6).cpu().numpy()
    assert np.abs(got - g["sfi_gray_mono"]).max() <= 4e-6 * g["sfi_gray_mono"].max()
    og = np.load(Path(__file__).parent / "golden" / "og_beat.npz")["rgb"]
    for stereo in (False, True):
        want = ao.spectrogram_from_image_array(og, power=0.25, stereo=stereo, max_value=30e6)
        got = image_util.spectrogram_from_image_device(torch.from_numpy(og).cuda(), power=0.25, stereo=stereo,
                                                       max_value=30e6).cpu().numpy()
        assert got.shape == want.shape == ((2 if stereo else 1), 512, 512)
        assert np.abs(got - want).max() <= 4e-6 * want.max()
    # a-6: device quantiser vs the reference's own output on the same amplitudes
    for key, spec in (("ifs_img2", g["ifs_spec2"]), ("ifs_img1", g["ifs_spec2"][:1])):
        img, mx = image_util.image_from_spectrogram_device(torch.from_numpy(spec.copy()).cuda(), power=0.25)
        d = np.abs(img.cpu().numpy().astype(np.int16) - g[key].astype(np.int16))
        assert float(mx) == float(spec.max())
        assert d.max() <= 1 and (d != 0).mean() < 2e-3, key             # truncation boundaries of powf


# ----------------------------------------------------------------------------------------------- generate_clips
def _chain_checks(out, lat_ref32, lat_emul, oracle_vae, conv, angles, what, lat_emul2=None):
    """stages after the loop, each re-synchronised on OUR previous stage so one stage is judged at a time"""
    from oracle import audio_oracle as ao
    from oracle.torchaudio_ref import TorchaudioConverter
    from oracle.vae_oracle import u8_from_image_fp16

    _check_vs_floor(out["latents_unscaled"], lat_ref32, lat_emul, what + ": loop latents", emul2=lat_emul2, floor_factor=1.25)
    # VAE decode of OUR latents by the fp32 oracle -> image, uint8
    zs = out["latents"]                                                # already 1/0.18215-scaled fp16 (reference :427)
    img32 = torch.cat([oracle_vae.decode(zs[i:i + 1].float()) for i in range(zs.shape[0])])
    u8_ref = u8_from_image_fp16(img32.half())
    u8 = out["images"].cpu().numpy()
    d = np.abs(u8.astype(np.int16) - u8_ref.astype(np.int16))
    print(f"{what}: uint8 image vs oracle VAE on the same latents: max |diff| {d.max()} LSB, differing pixels {100 * (d != 0).mean():.2f} %")
    assert d.max() <= 2 and (d != 0).mean() < 0.30 and (d > 1).mean() < 2e-3
    # image -> mel on OUR uint8 image (mono = R plane): exact arithmetic, powf rounding only
    B = u8.shape[0]
    mel_ref = np.concatenate([ao.spectrogram_from_image_array(u8[i], power=0.25, stereo=False, max_value=30e6) for i in range(B)])
    ta = TorchaudioConverter()
    wave_ref = ta.waveform_from_mel_amplitudes(torch.from_numpy(mel_ref), angles.cpu())
    wave = out["waveform"].cpu()
    assert wave.shape == wave_ref.shape == (B, 441 * 511)

    def nrms(a, b):
        return float((((a - b) / b.abs().amax(dim=-1, keepdim=True)) ** 2).mean().sqrt())

    rms = nrms(wave, wave_ref)
    print(f"{what}: waveform vs torchaudio inverse-mel + Griffin-Lim 32 it on our uint8 image: normalised RMS {rms:.3e}")
    if rms >= 1e-4:
        # Griffin-Lim amplifies fp32 rounding on ill-conditioned inputs (a random-weight VAE image is nearly flat): the
        # fp64 oracle recurrence is the referee, as in tests/test_audio_gpu.py — we must be no further from it than
        # twice torchaudio's own fp32 distance
        from riffusion.spectrogram_converter import mel_filterbank

        fb = mel_filterbank(8821, 0.0, 10000.0, 512, 44100).numpy()
        o64 = torch.from_numpy(ao.waveform_from_mel_amplitudes(mel_ref[:1], fb, 17640, 441, ao.hann_window(4410).double().numpy(),
                                                               32, angles[:1].cpu().numpy())).float()
        e_ta, e_us = nrms(wave_ref[:1], o64), nrms(wave[:1], o64)
        print(f"{what}: vs the fp64 oracle recurrence (clip 0): torchaudio fp32 {e_ta:.3e}, kernels {e_us:.3e}")
        assert e_us <= max(2 * e_ta, 2e-5)
    # un-synchronised end to end (oracle loop -> oracle VAE -> uint8): reported, loosely bounded (chaotic amplification
    # of the fp16 floor through the loop, then quantisation)
    img_e2e = torch.cat([oracle_vae.decode((lat_ref32[i:i + 1] / 0.18215)) for i in range(B)])
    d2 = np.abs(u8.astype(np.int16) - u8_from_image_fp16(img_e2e.half()).astype(np.int16))
    print(f"{what}: uint8 image vs the fp32 oracle chain end to end: mean |diff| {d2.mean():.3f} LSB, max {d2.max()}")
    assert d2.mean() < 1.0


@torch.no_grad()
def test_generate_clips_full_size_8_evals(sd15, vae_pair, conv):
    """the benchmarked call (`generate_clips`: denoise -> VAE decode -> uint8 -> mel -> inverse mel + 32-it Griffin-Lim) at
    full size, 2 clips x 8 CFG evaluations, against uo.img2img_loop -> AutoencoderKLOracle.decode -> numpy_to_pil rounding
    -> ao.spectrogram_from_image_array -> torchaudio with the same injected phases
    (riffusion_pipeline.py:398-434 + server.py:145-164)"""
    from oracle import unet_emul as ue
    from oracle import unet_oracle as uo
    from riffusion.riffusion_pipeline import RiffusionPipeline

    oracle, unet = sd15
    oracle_vae, vae = vae_pair
    pipe = RiffusionPipeline(vae=vae, unet=unet, device="cuda")
    torch.manual_seed(88)
    B = 2
    lat = (torch.randn(B, 4, 64, 64, device="cuda") * 0.18215 * 3).half()
    noise = torch.randn(B, 4, 64, 64, device="cuda").half()
    text = torch.randn(B, 77, 768, device="cuda").half()
    uncond = torch.randn(1, 77, 768, device="cuda").half()
    angles = torch.rand(B, 8821, 512, dtype=torch.complex64, device="cuda")
    out = pipe.generate_clips(text, uncond, lat, noise, 1.0, 8, 7.0, conv, init_angles=angles)
    assert out["n_unet_evals"] == 8 and out["images"].shape == (B, 512, 512, 3)
    refs, emuls = [], []
    for i in range(B):
        r, n = uo.img2img_loop(oracle, uo.PNDMSchedulerOracle(), text[i:i + 1].float(), uncond.float(), lat[i:i + 1].float(),
                               noise[i:i + 1].float(), noise[i:i + 1].float(), 0.0, 1.0, 8, 7.0)
        assert n == 8
        refs.append(r)
        emuls.append(ue.img2img_loop_emul(oracle, uo.PNDMSchedulerOracle(), text[i:i + 1], uncond, lat[i:i + 1], noise[i:i + 1],
                                          1.0, 8, 7.0)[0])
    emul2 = _emul_hi(ue.img2img_loop_emul, oracle, uo.PNDMSchedulerOracle(), text[0:1], uncond, lat[0:1], noise[0:1], 1.0, 8, 7.0)[0]
    print(f"8-eval loop, two fp16-storage evaluations apart (clip 0): {rel_l2(emul2, emuls[0]):.3e}")
    _chain_checks(out, torch.cat(refs), torch.cat(emuls), oracle_vae, conv, angles, "generate_clips full size, 8 evals")


@torch.no_grad()
def test_generate_clips_50_evals_small_unet(small_unet, vae_pair, conv):
    """the full 50-evaluation schedule (denoising 1.0) on a reduced-width UNet at 64x64 latents + the full-size VAE"""
    from oracle import unet_emul as ue
    from oracle import unet_oracle as uo
    from riffusion.riffusion_pipeline import RiffusionPipeline

    oracle, unet = small_unet
    oracle_vae, vae = vae_pair
    pipe = RiffusionPipeline(vae=vae, unet=unet, device="cuda")
    torch.manual_seed(50)
    lat = (torch.randn(1, 4, 64, 64, device="cuda") * 0.18215 * 3).half()
    noise = torch.randn(1, 4, 64, 64, device="cuda").half()
    text = torch.randn(1, 77, 64, device="cuda").half()
    uncond = torch.randn(1, 77, 64, device="cuda").half()
    angles = torch.rand(1, 8821, 512, dtype=torch.complex64, device="cuda")
    out = pipe.generate_clips(text, uncond, lat, noise, 1.0, 50, 7.0, conv, init_angles=angles)
    assert out["n_unet_evals"] == 50
    ref, n = uo.img2img_loop(oracle, uo.PNDMSchedulerOracle(), text.float(), uncond.float(), lat.float(), noise.float(),
                             noise.float(), 0.0, 1.0, 50, 7.0)
    assert n == 50
    emul, _ = ue.img2img_loop_emul(oracle, uo.PNDMSchedulerOracle(), text, uncond, lat, noise, 1.0, 50, 7.0)
    emul2 = _emul_hi(ue.img2img_loop_emul, oracle, uo.PNDMSchedulerOracle(), text, uncond, lat, noise, 1.0, 50, 7.0)[0]
    _chain_checks(out, ref, emul, oracle_vae, conv, angles, "generate_clips small UNet, 50 evals", lat_emul2=emul2)


# ----------------------------------------------------------------------------------------------- riffuse()
def _StubTokenizer():
    """stand-in for CLIPTokenizer (the text encoder is outside rows b-1..b-6): tests/golden/prompt_stub.py"""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).parent / "golden"))
    from prompt_stub import StubTokenizer

    return StubTokenizer()


class _StubTextEncoder:
    def __init__(self, dim=768):
        g = torch.Generator().manual_seed(123)
        self.table = torch.randn(49408, dim, generator=g).cuda()
        self.pos = torch.randn(77, dim, generator=g).cuda() * 0.3

    def __call__(self, ids):
        return ((self.table[ids.cuda()] + self.pos[None]).half(),)


@torch.no_grad()
def test_riffuse_pil_to_pil_with_mask(sd15, vae_pair):
    """BASELINE configs[2]: RiffusionPipeline.riffuse (riffusion_pipeline.py:208-287) on seed image og_beat, alpha 0.5,
    50 scheduler steps, denoising 0.75 (-> 38 CFG evaluations), guidance 7, WITH a mask (:420-425, preprocess_mask
    :455-477), PIL in -> PIL out, against the oracle executing the same steps with the same generator draws."""
    from pathlib import Path

    from PIL import Image

    from oracle import unet_emul as ue
    from oracle import unet_oracle as uo
    from oracle.vae_oracle import u8_from_image_fp16
    from riffusion.datatypes import InferenceInput, PromptInput
    from riffusion.riffusion_pipeline import RiffusionPipeline, preprocess_image, preprocess_mask

    oracle, unet = sd15
    oracle_vae, vae = vae_pair
    pipe = RiffusionPipeline(vae=vae, unet=unet, text_encoder=_StubTextEncoder(), tokenizer=_StubTokenizer(), device="cuda")
    pipe.device_slerp = False          # the reference's host-numpy fp16 slerp, so both sides see the same noise bit for bit
    rgb = np.load(Path(__file__).parent / "golden" / "og_beat.npz")["rgb"]
    init_image = Image.fromarray(rgb, mode="RGB")
    m = np.zeros((512, 512), dtype=np.uint8)
    m[:170] = 255                                                       # repaint the top third (like seed_images/mask_top_third_*.png)
    m[170:200] = 128
    mask_image = Image.fromarray(m, mode="L")
    inputs = InferenceInput(alpha=0.5, num_inference_steps=50, seed_image_id="og_beat",
                            start=PromptInput(prompt="church bells on sunday", seed=42),
                            end=PromptInput(prompt="jazz with piano", seed=123))
    for mask_img in (mask_image, None):
        got_img = pipe.riffuse(inputs, init_image, mask_img, use_reweighting=False)
        assert isinstance(got_img, Image.Image) and got_img.size == (512, 512) and got_img.mode == "RGB"
        # ---- oracle, same control flow (:227-287) in fp32
        alpha = inputs.alpha
        guidance = inputs.start.guidance * (1 - alpha) + inputs.end.guidance * alpha
        gen_a = torch.Generator(device="cuda").manual_seed(inputs.start.seed)
        gen_b = torch.Generator(device="cuda").manual_seed(inputs.end.seed)
        e0, e1 = pipe.embed_text(inputs.start.prompt), pipe.embed_text(inputs.end.prompt)
        text = (e0 + alpha * (e1 - e0))
        uncond = pipe.embed_text("")
        img_t = preprocess_image(init_image).cuda()
        mean, logvar = oracle_vae.encode_moments(img_t.half().float())
        gen = torch.Generator(device="cuda").manual_seed(inputs.start.seed)
        eps = torch.randn(mean.shape, generator=gen, device="cuda")
        init_latents = 0.18215 * (mean + torch.exp(0.5 * logvar) * eps)
        mask = None if mask_img is None else preprocess_mask(mask_img, 8).cuda()
        na = torch.randn(init_latents.shape, generator=gen_a, device="cuda", dtype=torch.float16)
        nb = torch.randn(init_latents.shape, generator=gen_b, device="cuda", dtype=torch.float16)
        strength = (1 - alpha) * inputs.start.denoising + alpha * inputs.end.denoising
        noise = uo.slerp(alpha, na, nb)                                  # fp16 numpy, like the reference
        ref, n = uo.img2img_loop(oracle, uo.PNDMSchedulerOracle(), text.float(), uncond.float(), init_latents, noise.float(),
                                 noise.float(), 0.0, strength, 50, guidance, mask=mask)
        assert n == 38
        emul, _ = ue.img2img_loop_emul(oracle, uo.PNDMSchedulerOracle(), text, uncond, init_latents.half(), noise, strength, 50,
                                       guidance, mask=mask)
        # our latents for the same request (the tap the reference exposes: dict["latents"], :436)
        lat0 = pipe.encode_image(init_image, torch.Generator(device="cuda").manual_seed(inputs.start.seed))
        e_enc = rel_l2(lat0, init_latents)
        out = pipe.interpolate_img2img(text_embeddings=text, init_latents=lat0, mask=None if mask is None else mask.half(),
                                       generator_a=torch.Generator(device="cuda").manual_seed(inputs.start.seed),
                                       generator_b=torch.Generator(device="cuda").manual_seed(inputs.end.seed),
                                       interpolate_alpha=alpha, strength_a=inputs.start.denoising, strength_b=inputs.end.denoising,
                                       num_inference_steps=50, guidance_scale=guidance, output_type="latent")
        assert out["n_unet_evals"] == 38
        e_k, e_f, e_o = rel_l2(out["latents_unscaled"], emul), rel_l2(emul, ref), rel_l2(out["latents_unscaled"], ref)
        tag = "mask" if mask_img is not None else "no mask"
        print(f"riffuse ({tag}): VAE-encode sample vs oracle {e_enc:.3e}; 38-eval loop latents: kernels vs emulated loop {e_k:.3e} | "
              f"fp16 floor of the loop {e_f:.3e} | kernels vs fp32 loop {e_o:.3e}")
        assert e_enc < 2e-3
        assert e_o <= 1.25 * e_f + 2e-4 and e_k <= 1.25 * (2 ** 0.5 * 1.25 * e_f) + 2e-4
        ref_u8 = u8_from_image_fp16(oracle_vae.decode(ref / 0.18215).half())[0]
        d = np.abs(np.array(got_img).astype(np.int16) - ref_u8.astype(np.int16))
        print(f"riffuse ({tag}): PIL image vs fp32 oracle chain: mean |diff| {d.mean():.3f} LSB, max {d.max()}, "
              f"pixels within 1 LSB {100 * (d <= 1).mean():.1f} %")
        assert d.mean() < 1.5
        if mask is not None:
            # the kept (mask = 1 after inversion -> black in the mask image) region follows the re-noised original latents:
            # at the last step t = 1 it is init_latents up to sqrt(1 - a_1) noise, identical on both sides
            keep = (mask[0, 0] > 0.99)
            assert keep.any()
            assert rel_l2(out["latents_unscaled"][0][:, keep], ref[0][:, keep]) < 2e-3


@torch.no_grad()
def test_riffuse_batch_equals_single_requests(small_unet, vae_pair):
    """`riffuse_batch` runs a list of InferenceInput in one batched loop; request i must reproduce
    `riffuse(inputs[i])` (same generator draws, per-request alpha) up to batch-size dependent accumulation order."""
    from pathlib import Path

    from PIL import Image

    from riffusion.datatypes import InferenceInput, PromptInput
    from riffusion.riffusion_pipeline import RiffusionPipeline

    _, unet = small_unet
    _, vae = vae_pair
    pipe = RiffusionPipeline(vae=vae, unet=unet, text_encoder=_StubTextEncoder(dim=64), tokenizer=_StubTokenizer(), device="cuda")
    rgb = np.load(Path(__file__).parent / "golden" / "og_beat.npz")["rgb"]
    init_image = Image.fromarray(rgb, mode="RGB")
    reqs = [InferenceInput(alpha=a, num_inference_steps=20, seed_image_id="og_beat",
                           start=PromptInput(prompt="church bells on sunday", seed=42 + i, denoising=0.75, guidance=7.0),
                           end=PromptInput(prompt="jazz with (piano:1.2)", seed=123 + i, denoising=0.75, guidance=7.0))
            for i, a in enumerate((0.0, 0.3, 1.0))]
    reqs.append(InferenceInput(alpha=0.5, num_inference_steps=20, seed_image_id="og_beat",            # its own group
                               start=PromptInput(prompt="a", seed=1, denoising=0.5), end=PromptInput(prompt="b", seed=2, denoising=0.5)))
    batch = pipe.riffuse_batch(reqs, init_image)
    assert len(batch) == 4 and all(im.size == (512, 512) for im in batch)
    for i, r in enumerate(reqs):
        single = pipe.riffuse(r, init_image)
        d = np.abs(np.array(batch[i]).astype(np.int16) - np.array(single).astype(np.int16))
        print(f"riffuse_batch request {i}: vs riffuse() mean |diff| {d.mean():.4f} LSB, max {d.max()}, equal pixels {100 * (d == 0).mean():.1f} %")
        assert d.mean() < 0.25 and (d <= 1).mean() > 0.98
    assert np.abs(np.array(batch[0]).astype(np.int16) - np.array(batch[2]).astype(np.int16)).mean() > 0.5   # requests differ
