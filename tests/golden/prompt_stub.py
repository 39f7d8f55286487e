"""Deterministic stand-ins for CLIPTokenizer / CLIPTextModel, shared by tests/golden/make_golden_prompt.py (which feeds
them to the REFERENCE's prompt-weighting functions) and tests/test_prompt_weighting_cpu.py (which feeds them to ours)."""
import types
import zlib

import torch


class StubTokenizer:
    model_max_length = 77
    bos_token_id = 49406
    eos_token_id = 49407

    def _ids(self, text: str):
        words = text.replace(",", " , ").replace(".", " . ").split()
        return [self.bos_token_id] + [1 + zlib.crc32(w.lower().encode()) % 49000 for w in words] + [self.eos_token_id]

    def __call__(self, text, padding=None, max_length=None, truncation=False, return_tensors=None):
        single = isinstance(text, str)
        rows = [self._ids(t) for t in ([text] if single else text)]
        if truncation and max_length:
            rows = [r[: max_length - 1] + [self.eos_token_id] if len(r) > max_length else r for r in rows]
        if padding == "max_length":
            rows = [r + [self.eos_token_id] * (max_length - len(r)) for r in rows]
        if return_tensors == "pt":
            return types.SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.long))
        return types.SimpleNamespace(input_ids=rows[0] if single else rows)


class StubTextEncoder:
    """embedding[token] + position, mixed through one fixed linear layer: (B, 77) ids -> ((B, 77, 32) fp32,)"""

    def __init__(self, dim: int = 32):
        g = torch.Generator().manual_seed(2024)
        self.table = torch.randn(49408, dim, generator=g)
        self.pos = torch.randn(77, dim, generator=g)
        self.mix = torch.randn(dim, dim, generator=g) / dim ** 0.5

    def __call__(self, ids):
        assert ids.shape[1] == 77, ids.shape
        return ((self.table[ids] + self.pos[None]) @ self.mix + 0.1,)


def stub_pipe():
    return types.SimpleNamespace(tokenizer=StubTokenizer(), text_encoder=StubTextEncoder(), device="cpu")
