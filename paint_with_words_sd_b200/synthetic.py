"""Stand-ins for the assets the reference downloads (CLIP vocab, CLIP text encoder, VAE).

There is no network, no model weights and no CLIP vocabulary in the build or bench environment
(SURVEY.md headline facts), so end-to-end runs are synthetic: seeded random weights, a deterministic
word-hash tokenizer with the CLIPTokenizer call protocol, and a random-projection text encoder with the
CLIPTextModel call protocol.  Real `transformers` objects can be passed instead through
`preloaded_utils`, exactly like the reference (paint_with_words.py:408, 415-425).
"""
from __future__ import annotations

import re
import zlib
from typing import List, Sequence, Union

import torch


class _Encoding(dict):
    """dict with attribute access, like transformers.BatchEncoding (`enc.input_ids`, `enc["input_ids"]`)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class SimpleWordTokenizer:
    """Deterministic word-level tokenizer with the CLIPTokenizer call protocol used at
    paint_with_words.py:222-227 and 320-326: BOS=49406, EOS=49407, padding with EOS to
    `model_max_length`=77, ids of words are a CRC32 hash into [1000, 49000)."""

    bos_token_id = 49406
    eos_token_id = 49407
    model_max_length = 77
    _pat = re.compile(r"[a-z0-9]+|[^\sa-z0-9]")

    def _ids(self, text: str) -> List[int]:
        return [1000 + zlib.crc32(w.encode()) % 48000 for w in self._pat.findall(text.lower())]

    def __call__(self, text: Union[str, Sequence[str]], padding=None, max_length=None, truncation=False,
                 return_tensors=None, **_):
        batched = not isinstance(text, str)
        texts = list(text) if batched else [text]
        max_length = max_length or self.model_max_length
        rows = []
        for t in texts:
            ids = self._ids(t)
            if truncation:
                ids = ids[: max_length - 2]
            ids = [self.bos_token_id] + ids + [self.eos_token_id]
            if padding == "max_length":
                ids = ids + [self.eos_token_id] * (max_length - len(ids))
            rows.append(ids)
        if return_tensors == "pt":
            return _Encoding(input_ids=torch.tensor(rows, dtype=torch.long))
        return _Encoding(input_ids=rows if batched else rows[0])


class RandomTextEncoder(torch.nn.Module):
    """CLIPTextModel stand-in: `enc(input_ids)[0]` -> [B,77,dim] fp32 (the reference keeps the text
    encoder in fp32, paint_with_words.py:171).  Embedding table + positional table, seeded."""

    def __init__(self, dim: int = 768, vocab: int = 49408, max_len: int = 77, seed: int = 1234):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.tok = torch.nn.Parameter(torch.randn(vocab, dim, generator=g), requires_grad=False)
        self.pos = torch.nn.Parameter(0.1 * torch.randn(max_len, dim, generator=g), requires_grad=False)

    def forward(self, input_ids):
        x = self.tok[input_ids] + self.pos[None, : input_ids.shape[1]]
        return (torch.nn.functional.layer_norm(x, x.shape[-1:]),)


class IdentityVAE(torch.nn.Module):
    """AutoencoderKL stand-in: decode(latents).sample upsamples the first 3 latent channels x8 (the
    VAE is outside the hot path; it exists so `paint_with_words()` can return a PIL image)."""

    class _Out:
        def __init__(self, s):
            self.sample = s

    class _Dist:
        def __init__(self, s):
            self._s = s

        def sample(self):
            return self._s

    class _Enc:
        def __init__(self, s):
            self.latent_dist = IdentityVAE._Dist(s)

    def encode(self, image):
        x = torch.nn.functional.avg_pool2d(image.float(), 8)
        return self._Enc(torch.cat([x, x[:, :1]], dim=1))

    def decode(self, latents):
        img = torch.nn.functional.interpolate(latents[:, :3].float(), scale_factor=8, mode="nearest")
        return self._Out(torch.tanh(img))
