"""xformers.ops.memory_efficient_attention [3P restatement] — call site models/modules/transformer.py:71.
3-D inputs [B*H, M, K]: softmax(q k^T / sqrt(K) + bias) v (the SDPA form in the reference's own comment,
transformer.py:69-70)."""
import torch


def memory_efficient_attention(q, k, v, attn_bias=None, p=0.0, scale=None):
    scale = q.shape[-1] ** -0.5 if scale is None else scale
    s = torch.baddbmm(attn_bias, q, k.transpose(1, 2), alpha=scale) if attn_bias is not None \
        else torch.bmm(q, k.transpose(1, 2)) * scale
    return torch.bmm(torch.softmax(s, dim=-1), v)
