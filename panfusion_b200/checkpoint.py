"""Loading the reference's Lightning checkpoints (`last.ckpt`) into the drop-in `MultiViewBaseModel`.

What the reference does at load time (models/pano/PanoGenerator.py:84-114): `torch.load(ckpt)['state_dict']`,
`convert_state_dict` renames the LoRA keys between the two spellings diffusers uses
(`...to_q.lora_layer.{down,up}.weight` <-> `...processor.to_q_lora.{down,up}.weight`), `exclude_eval_metrics` drops
`eval_metrics*`, then `load_state_dict(strict=True)` with a non-strict retry. Because the UNets are wrapped by
`torch.compile` (PanoGenerator.py:176) their keys carry `_orig_mod.`.

Here the rank-4 LoRA adapters on every attention `to_q / to_k / to_v / to_out.0` (PanoGenerator.py:132-151; diffusers
`LoRALinearLayer` [3P]: `W x + scale * up(down(x))`, scale 1, no network_alpha) are FOLDED into the base weights,
`W' = W + up @ down`, because the kernels run plain GEMMs on packed weights. This is load-time weight preparation on
the host (like engine.UNetPack), not part of the compute path.
"""
from __future__ import annotations

import re
from typing import Mapping, Optional

import torch
from torch import Tensor

_LORA_A = re.compile(r"^(?P<base>.*)\.(?P<proj>to_q|to_k|to_v|to_out\.0)\.lora_layer\.(?P<part>down|up)\.weight$")
_LORA_B = re.compile(r"^(?P<base>.*)\.processor\.(?P<proj>to_q|to_k|to_v|to_out)_lora\.(?P<part>down|up)\.weight$")


def split_reference_state_dict(state_dict: Mapping[str, Tensor], prefix: str = "mv_base_model."):
    """-> (plain, lora): `plain` maps MultiViewBaseModel parameter names to tensors (prefix and `_orig_mod.` removed,
    keys outside `prefix` — VAE, text encoder, eval metrics — dropped); `lora` maps the name of an attention Linear
    weight (e.g. `unet.down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.weight`) to {'down': [r, in],
    'up': [out, r]}. Both LoRA key spellings are understood (PanoGenerator.py:101-107)."""
    plain, lora = {}, {}
    for key, value in state_dict.items():
        if not key.startswith(prefix):
            continue
        name = key[len(prefix):].replace("_orig_mod.", "")
        m = _LORA_A.match(name) or _LORA_B.match(name)
        if m is None:
            plain[name] = value
            continue
        proj = m.group("proj")
        if proj == "to_out":
            proj = "to_out.0"
        lora.setdefault(f"{m.group('base')}.{proj}.weight", {})[m.group("part")] = value
    for target, parts in lora.items():
        if set(parts) != {"down", "up"}:
            raise KeyError(f"incomplete LoRA pair for {target}: {sorted(parts)}")
    return plain, lora


def fold_lora(weight: Tensor, down: Tensor, up: Tensor, scale: float = 1.0) -> Tensor:
    """W' = W + scale * up @ down in fp32 (diffusers LoRACompatibleLinear: W x + scale * up(down(x)))."""
    return (weight.to(torch.float32) + scale * (up.to(torch.float32) @ down.to(torch.float32))).to(weight.dtype)


@torch.no_grad()
def load_reference_state_dict(model, state_dict: Mapping[str, Tensor], prefix: str = "mv_base_model.",
                              lora_scale: float = 1.0, strict: bool = False) -> dict:
    """Load a reference checkpoint's `state_dict` into `model` (panfusion_b200.mvgen.MultiViewBaseModel): plain
    tensors are copied by name, LoRA pairs are folded into the (just loaded, or already present) base weight. Returns a
    report {'loaded', 'folded', 'missing', 'unexpected'}; with strict=True missing / unexpected names raise like
    `nn.Module.load_state_dict`. The packed kernel weights are invalidated (re-packed on the next forward)."""
    plain, lora = split_reference_state_dict(state_dict, prefix)
    # the user's own UNets may be torch.compile'd as well: map clean names to the model's tensors
    own = {k.replace("_orig_mod.", ""): v for k, v in model.state_dict().items()}
    unexpected = sorted(k for k in plain if k not in own)
    missing = sorted(k for k in own if k not in plain)
    if not any(k in own for k in plain) and not any(t in own for t in lora):
        # nothing matched: wrong prefix (e.g. a state_dict saved without 'mv_base_model.') — never a silent no-op
        raise KeyError(f"no key of the checkpoint matches the model under prefix {prefix!r} "
                       f"(first keys: {list(state_dict)[:3]})")
    if lora and not plain:
        # a LoRA-only checkpoint folds into whatever base weights the model holds: folding it twice is silently wrong
        if getattr(model, "_lora_folded", False):
            raise RuntimeError("LoRA adapters were already folded into this model's weights; reload the base weights "
                               "before loading a LoRA-only checkpoint again")
    for name, value in plain.items():
        if name in own:
            if own[name].shape != value.shape:
                raise RuntimeError(f"size mismatch for {name}: checkpoint {tuple(value.shape)} vs model {tuple(own[name].shape)}")
            own[name].copy_(value)
    folded = 0
    for target, parts in lora.items():
        if target not in own:
            unexpected.append(target + " (LoRA)")
            continue
        own[target].copy_(fold_lora(own[target], parts["down"].to(own[target].device), parts["up"].to(own[target].device),
                                    lora_scale))
        folded += 1
    if lora:
        model._lora_folded = True
    non_lora_missing = [k for k in missing if "cp_blocks" in k]
    if non_lora_missing and not strict:
        import warnings
        warnings.warn(f"{len(non_lora_missing)} EPPA (cp_blocks) tensors are not in the checkpoint and keep their current "
                      f"values, e.g. {non_lora_missing[:3]}", stacklevel=2)
    if strict and (missing or unexpected):
        raise RuntimeError(f"missing keys: {missing[:8]}{'...' if len(missing) > 8 else ''}; "
                           f"unexpected keys: {unexpected[:8]}{'...' if len(unexpected) > 8 else ''}")
    if hasattr(model, "invalidate"):
        model.invalidate()
    return dict(loaded=len(plain) - len([k for k in plain if k not in own]), folded=folded, missing=missing,
                unexpected=sorted(unexpected))


def load_reference_checkpoint(model, path: str, map_location="cpu", allow_pickle: bool = False, **kw) -> dict:
    """`torch.load(path)['state_dict']` (PanoGenerator.py:88) -> load_reference_state_dict. Only tensors are needed, so
    the file is read with `weights_only=True`; Lightning checkpoints that carry pickled hyper-parameter objects need
    `allow_pickle=True` (arbitrary code execution from an untrusted file — opt in explicitly)."""
    try:
        ckpt = torch.load(path, map_location=map_location, weights_only=True)
    except Exception:
        if not allow_pickle:
            raise
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
    return load_reference_state_dict(model, ckpt["state_dict"] if "state_dict" in ckpt else ckpt, **kw)
