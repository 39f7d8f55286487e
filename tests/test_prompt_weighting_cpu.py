"""(f)-3 host part: the prompt attention-weighting parser and the weighted-embedding assembly against vectors produced by
the REFERENCE's own functions (tests/golden/make_golden_prompt.py imports riffusion/external/prompt_weighting.py) on
deterministic tokenizer / text-encoder stand-ins, plus the known answers printed in the reference's docstring (:43-75)."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden"
sys.path.insert(0, str(GOLD))


@pytest.fixture(scope="module")
def gold():
    return json.loads((GOLD / "prompt_vectors.json").read_text()), np.load(GOLD / "prompt_vectors.npz")


def test_parse_prompt_attention_docstring_known_answers():
    from riffusion.external.prompt_weighting import parse_prompt_attention as parse

    assert parse("normal text") == [["normal text", 1.0]]
    assert parse("an (important) word") == [["an ", 1.0], ["important", 1.1], [" word", 1.0]]
    assert parse("(unbalanced") == [["unbalanced", 1.1]]
    assert parse("\\(literal\\]") == [["(literal]", 1.0]]
    assert parse("(unnecessary)(parens)") == [["unnecessaryparens", 1.1]]
    got = parse("a (((house:1.3)) [on] a (hill:0.5), sun, (((sky))).")
    want = [["a ", 1.0], ["house", 1.5730000000000004], [" ", 1.1], ["on", 1.0], [" a ", 1.1], ["hill", 0.55],
            [", sun, ", 1.1], ["sky", 1.4641000000000006], [".", 1.1]]
    assert [g[0] for g in got] == [w[0] for w in want]
    assert np.allclose([g[1] for g in got], [w[1] for w in want], rtol=1e-12)


def test_parse_matches_reference_on_all_vectors(gold):
    from riffusion.external.prompt_weighting import parse_prompt_attention as parse

    js, _ = gold
    assert len(js["parsed"]) >= 20
    for prompt, want in js["parsed"].items():
        got = parse(prompt)
        assert [g[0] for g in got] == [w[0] for w in want], prompt
        assert [g[1] for g in got] == [w[1] for w in want], prompt        # same multiplication order -> identical floats


def test_tokens_weights_and_embeddings_match_reference(gold):
    from prompt_stub import stub_pipe
    from riffusion.external import prompt_weighting as pw

    js, npz = gold
    pipe = stub_pipe()
    toks, wts = pw.get_prompts_with_weights(pipe, [js["long"], "a (b:2) c"], 225)
    assert toks == js["tokens"] and wts == js["weights"]
    for name, prompt in js["cases"].items():
        e, u = pw.get_weighted_text_embeddings(pipe=pipe, prompt=prompt, uncond_prompt=None, max_embeddings_multiples=3,
                                               no_boseos_middle=False, skip_parsing=False, skip_weighting=False)
        want = npz["emb_" + name]
        assert u is None and tuple(e.shape) == want.shape, name
        assert np.allclose(e.numpy(), want, rtol=1e-6, atol=1e-6), name
    assert npz["emb_long"].shape == (1, 154, 32)          # 120 words -> two 77-row chunks (Appendix A-11)
    e, u = pw.get_weighted_text_embeddings(pipe=pipe, prompt=["(a:1.3) b", "c [d]"], uncond_prompt=["", "(e)"],
                                           max_embeddings_multiples=3, no_boseos_middle=True)
    assert np.allclose(e.numpy(), npz["emb_pair"], rtol=1e-6, atol=1e-6)
    assert np.allclose(u.numpy(), npz["unc_pair"], rtol=1e-6, atol=1e-6)


def test_pipeline_embed_text_weighted_routes_through_the_parser():
    """RiffusionPipeline.embed_text_weighted (riffusion_pipeline.py:193-206): parenthesised prompts are accepted,
    un-weighted prompts equal embed_text, weighted ones preserve the mean"""
    from prompt_stub import StubTextEncoder, StubTokenizer
    from riffusion.riffusion_pipeline import RiffusionPipeline

    pipe = RiffusionPipeline(vae=None, unet=None, text_encoder=StubTextEncoder(), tokenizer=StubTokenizer(), device="cpu")
    plain = pipe.embed_text_weighted("jazz live")
    assert torch.allclose(plain.float(), pipe.embed_text("jazz live").float(), atol=2e-3)
    w = pipe.embed_text_weighted("jazz (live:1.4)")
    assert w.shape == plain.shape and not torch.allclose(w, plain)
    assert abs(float(w.float().mean()) - float(plain.float().mean())) < 2e-3
    assert pipe.embed_text_weighted("jazz (live)").shape == plain.shape       # used to raise
