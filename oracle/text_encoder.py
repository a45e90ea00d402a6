"""Oracle for the text encoder: transformers' own `CLIPTextModel` [3P] — the class the reference instantiates
(models/pano/PanoGenerator.py:117-121) — EXECUTED, not restated: the library is installed in this image (v5.5; the reference
pins 4.x, whose CLIP text tower computes the same function). SD-2-base `text_encoder/config.json` values are restated from
the published checkpoint (hidden 1024, 16 heads, 23 layers, intermediate 4096, gelu, eps 1e-5, 77 positions, vocab 49408);
weights are seeded random draws (no checkpoint offline). Test infrastructure only."""
from __future__ import annotations

import torch

SD2_TEXT_CONFIG = dict(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                       num_attention_heads=16, max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5,
                       projection_dim=512)
TINY_TEXT_CONFIG = dict(SD2_TEXT_CONFIG, vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=3,
                        num_attention_heads=2)


def build_text_encoder(config=None, seed: int = 0):
    from transformers import CLIPTextConfig, CLIPTextModel
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    model = CLIPTextModel(CLIPTextConfig(**(config or SD2_TEXT_CONFIG))).eval()
    with torch.no_grad():  # default init draws tiny weights (std 0.02 / scaled): widen so that every layer matters
        for name, p in model.named_parameters():
            if p.dim() == 2 and "embedding" not in name:
                p.mul_(4.0)
            elif p.dim() == 1 and "bias" in name:
                p.copy_(torch.randn(p.shape) * 0.05)
            elif p.dim() == 1:
                p.copy_(1.0 + 0.2 * torch.randn(p.shape))
    torch.random.set_rng_state(st)
    return model


def token_ids(batch: int, length: int = 77, vocab: int = 49408, seed: int = 0, eos: int = 49407):
    """Synthetic token rows shaped like the tokenizer's output: BOS, a few words, EOS padding to max_length."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.full((batch, length), min(eos, vocab - 1), dtype=torch.int64)
    ids[:, 0] = min(49406, vocab - 2)
    for b in range(batch):
        n = int(torch.randint(3, length - 2, (1,), generator=g))
        ids[b, 1:1 + n] = torch.randint(0, vocab - 2, (n,), generator=g)
    return ids


@torch.no_grad()
def encode_text(model, input_ids):
    """PanoGenerator.encode_text after the tokenizer (:207-211): `text_encoder(input_ids, attention_mask=None)[0]`."""
    return model(input_ids)[0]
