"""Golden vectors for prompt weighting ((f)-3), produced by IMPORTING the reference's own
riffusion/external/prompt_weighting.py (with a placeholder `diffusers` module: the file only uses it for a type
annotation) and running its functions on the deterministic tokenizer / text-encoder stand-ins of prompt_stub.py:

    python tests/golden/make_golden_prompt.py     # needs /root/reference; writes tests/golden/prompt_vectors.{json,npz}

Only inputs and numbers are stored — no reference source.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, "/root/reference")
fake = types.ModuleType("diffusers")
fake.StableDiffusionPipeline = object
sys.modules["diffusers"] = fake
from riffusion.external import prompt_weighting as ref  # noqa: E402  (the reference's)

assert ref.__file__.startswith("/root/reference")
from prompt_stub import stub_pipe  # noqa: E402

PROMPTS = [
    "normal text", "an (important) word", "(unbalanced", "\\(literal\\]", "(unnecessary)(parens)",
    "a (((house:1.3)) [on] a (hill:0.5), sun, (((sky))).", "", "jazz (live)", "[soft] piano, (loud:1.5) drums",
    "church bells on sunday", "time: 12:30) noon", "a ] b ) c", "((nested [mixed) case]", "back\\slash and \\\\ double",
    "(x:+2)(y:-0.5)(z:.5)", "[[[quiet]]] (((LOUD)))", "trailing colon:", "(a:1.1) (b:1.1)", "emoji ♫ (música:1.2)",
]
LONG = " ".join(f"(word{i}:1.{i % 9})" if i % 7 == 0 else f"word{i}" for i in range(120))


def main():
    parsed = {p: ref.parse_prompt_attention(p) for p in PROMPTS + [LONG]}
    pipe = stub_pipe()
    emb = {}
    cases = {"plain": "church bells on sunday", "weighted": "[soft] piano, (loud:1.5) drums, (((sky)))", "empty": "",
             "long": LONG}
    for name, prompt in cases.items():
        e, _ = ref.get_weighted_text_embeddings(pipe=pipe, prompt=prompt, uncond_prompt=None, max_embeddings_multiples=3,
                                                no_boseos_middle=False, skip_parsing=False, skip_weighting=False)
        emb["emb_" + name] = e.numpy()
    e, u = ref.get_weighted_text_embeddings(pipe=pipe, prompt=["(a:1.3) b", "c [d]"], uncond_prompt=["", "(e)"],
                                            max_embeddings_multiples=3, no_boseos_middle=True)
    emb["emb_pair"], emb["unc_pair"] = e.numpy(), u.numpy()
    toks, wts = ref.get_prompts_with_weights(pipe, [LONG, "a (b:2) c"], 225)
    (HERE / "prompt_vectors.json").write_text(json.dumps(
        {"parsed": parsed, "cases": cases, "long": LONG, "tokens": toks, "weights": wts}, ensure_ascii=False, indent=0))
    np.savez_compressed(HERE / "prompt_vectors.npz", **emb)
    print({k: v.shape for k, v in emb.items()})


if __name__ == "__main__":
    main()
