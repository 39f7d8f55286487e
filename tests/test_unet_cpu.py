"""CPU checks that pin the path (b) oracle as far as the environment allows (diffusers is not installable):
published parameter count, scheduler constants and the PLMS start-index table of SURVEY Appendix B."""
import torch

from oracle import unet_oracle as uo


def test_sd15_parameter_count_matches_published():
    with torch.device("meta"):
        m = uo.UNet2DConditionOracle()
    assert sum(p.numel() for p in m.parameters()) == 859_520_964     # "860M" UNet of Stable Diffusion 1.x


def test_pndm_tables_and_constants():
    s = uo.PNDMSchedulerOracle()
    assert abs(float(s.alphas_cumprod[0]) - 0.99915) < 1e-6 and abs(float(s.alphas_cumprod[999]) - 0.0046601) < 1e-6
    s.set_timesteps(50)
    ts = s.timesteps.tolist()
    assert len(ts) == 51 and ts[:4] == [981, 961, 961, 941] and ts[-2:] == [21, 1]
    # riffusion_pipeline.py:358-396 start arithmetic with steps_offset = 1
    for strength, t0, n_evals in ((0.75, 741, 38), (1.0, 961, 50), (0.5, 501, 26)):
        init = min(int(50 * strength) + 1, 50)
        assert ts[-init] == t0
        assert len(ts[max(50 - init + 1, 0):]) == n_evals
    assert abs(float(s.alphas_cumprod[741]) ** 0.5 - 0.245674) < 1e-5


def test_plms_step_sequence_on_a_linear_model():
    """the counter==1 re-step and the Adams-Bashforth weights: with a constant model output e the update
    x' = ca x - cb e must be applied exactly once per call (the 2nd call restarts from the saved sample)"""
    s = uo.PNDMSchedulerOracle()
    s.set_timesteps(50)
    x = torch.ones(1, 4, 2, 2)
    e = torch.full_like(x, 0.5)
    ts = s.timesteps[13:].tolist()
    x1 = s.step(e, ts[0], x)
    ca, cb = s.coefficients(741, 721)
    assert torch.allclose(x1, ca * x - cb * e)
    x2 = s.step(e, ts[1], x1)                    # counter == 1: redo 741 -> 721 from the saved sample
    assert torch.allclose(x2, ca * x - cb * e)
    x3 = s.step(e, ts[2], x2)
    ca3, cb3 = s.coefficients(701, 681)
    assert torch.allclose(x3, ca3 * x2 - cb3 * e)


def test_small_unet_runs_and_slerp_lerp_fallback():
    m = uo.init_weights_(uo.UNet2DConditionOracle(block_out_channels=(64, 128, 128, 128), heads=4, cross_attention_dim=64))
    y = m(torch.randn(2, 4, 16, 16), 741, torch.randn(2, 77, 64))
    assert y.shape == (2, 4, 16, 16) and torch.isfinite(y).all()
    a = torch.randn(16384)
    out = uo.slerp(0.3, a, a * 1.0001)           # |dot| > 0.9995 -> lerp (torch_util.py:34-35)
    assert torch.allclose(out, 0.7 * a + 0.3 * a * 1.0001, atol=1e-6)


def test_product_spec_matches_oracle_state_dicts():
    """riffusion.sd15_spec (product side, used for random init / checkpoint validation) enumerates exactly the
    parameters of the oracle modules"""
    from oracle.vae_oracle import AutoencoderKLOracle
    from riffusion import sd15_spec

    with torch.device("meta"):
        u, v = uo.UNet2DConditionOracle(), AutoencoderKLOracle()
    for spec, module in ((sd15_spec.unet_spec(), u), (sd15_spec.vae_spec(), v)):
        want = {k: tuple(p.shape) for k, p in module.state_dict().items()}
        got = dict(spec)
        assert len(got) == len(spec) and got == want
    assert sum(p.numel() for p in v.parameters()) == 83_653_863       # published SD VAE size


def test_geglu_weight_interleave_matches_chunk_semantics():
    """tc_ops.interleave_geglu reorders GEGLU.proj rows into runs of [16 value | 16 gate]; the fused epilogue of
    rf_gemm_f16 (act = 2) then computes out[:, 16 r + j] = v_j * gelu(g_j) from columns 32 r + j and 32 r + 16 + j.
    Emulated here in torch against diffusers' GEGLU: hidden, gate = proj(x).chunk(2, -1); hidden * gelu(gate)."""
    import torch.nn.functional as F

    from riffusion import tc_ops

    torch.manual_seed(0)
    C, inner, M = 48, 64, 10                       # inner % 16 == 0
    x = torch.randn(M, C)
    w, b = torch.randn(2 * inner, C), torch.randn(2 * inner)
    hidden, gate = (x @ w.t() + b).chunk(2, dim=-1)
    ref = hidden * F.gelu(gate)
    y = x @ tc_ops.interleave_geglu(w).t() + tc_ops.interleave_geglu(b)          # what the GEMM accumulates
    y = y.reshape(M, inner // 16, 2, 16)
    got = (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(M, inner)
    assert torch.allclose(got, ref, atol=1e-5)


def test_product_scheduler_bookkeeping_matches_oracle(monkeypatch):
    """host side of PNDMSchedulerB200 (timestep table, counter == 1 re-step, history rotation, Adams-Bashforth weights,
    ca / cb) against the oracle scheduler, with the fused device kernel (rf_cfg_pndm_step_f16) replaced by its torch
    definition: eps = eu + g (et - eu); e = sum c_i h_i; x' = ca x - cb e"""
    from riffusion import tc_ops
    from riffusion.scheduler_b200 import PNDMSchedulerB200

    def fake_step(eps_pair, guidance, hist, coef, sample, ca, cb, want_eps=True):
        n = sample.shape[0]
        eu, et = eps_pair[:n].double(), eps_pair[n:].double()
        eps = eu + guidance * (et - eu)
        e = coef[0] * eps
        for c, h in zip(coef[1:], hist):
            e = e + c * h.double()
        return (eps if want_eps else None), ca * sample.double() - cb * e

    monkeypatch.setattr(tc_ops, "cfg_pndm_step", fake_step)
    ours, ref = PNDMSchedulerB200(), uo.PNDMSchedulerOracle()
    ours.set_timesteps(50)
    ref.set_timesteps(50)
    assert [int(t) for t in ours.timesteps] == [int(t) for t in ref.timesteps]
    torch.manual_seed(3)
    x_o = x_r = torch.randn(2, 4, 8, 8, dtype=torch.float64)
    g = 7.0
    for t in [int(v) for v in ref.timesteps[13:24]]:                # 11 steps: Euler, re-step, 2-, 3-, 4-term PLMS
        pair = torch.randn(4, 4, 8, 8, dtype=torch.float64)
        guided = pair[:2] + g * (pair[2:] - pair[:2])
        x_r = ref.step(guided, t, x_r)
        x_o = ours.step_cfg(pair, g, t, x_o)
        assert torch.allclose(x_o, x_r, rtol=1e-5, atol=1e-5), t      # ca / cb are fp32 table look-ups on both sides
