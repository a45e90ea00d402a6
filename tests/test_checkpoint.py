"""CPU: loading a reference-style Lightning state_dict (models/pano/PanoGenerator.py:84-114) into the drop-in model:
`mv_base_model.` prefix, `_orig_mod.` from torch.compile, both LoRA key spellings folded as W + up @ down, foreign
keys (VAE, eval metrics) ignored."""
import pytest
import torch

from oracle import mvgen as om, synth, unet as ou
from panfusion_b200 import checkpoint as ck
from panfusion_b200.mvgen import MultiViewBaseModel


def _attention_linears(sd):
    return [k for k in sd if k.endswith(".weight") and any(
        k.endswith(f".{a}.{p}.weight") for a in ("attn1", "attn2") for p in ("to_q", "to_k", "to_v", "to_out.0"))]


def _reference_style(src, rank=4, seed=0):
    """state_dict the way the reference's checkpoints look + the expected folded weights."""
    g = torch.Generator().manual_seed(seed)
    sd, expect = {}, {}
    own = src.state_dict()
    for i, (name, w) in enumerate(own.items()):
        key = name
        for branch in ("unet.", "pano_unet."):
            if name.startswith(branch):
                key = branch + "_orig_mod." + name[len(branch):]  # torch.compile wrapper (PanoGenerator.py:176)
        sd["mv_base_model." + key] = w.clone()
        expect[name] = w.clone()
    for j, name in enumerate(n for n in _attention_linears(own) if n.startswith(("unet.", "pano_unet."))):
        w = own[name]
        down, up = torch.randn(rank, w.shape[1], generator=g) * 0.05, torch.randn(w.shape[0], rank, generator=g) * 0.05
        branch, rest = name.split(".", 1)
        proj_a = "to_out.0" if rest.endswith("to_out.0.weight") else rest.split(".")[-2]
        proj_b = proj_a.split(".")[0]                      # the processor spelling drops the ".0" of to_out
        base = rest[:-len(f".{proj_a}.weight")]
        stem = f"mv_base_model.{branch}._orig_mod.{base}"
        if j % 2 == 0:   # spelling written by on_save_checkpoint after the processors were swapped out
            sd[f"{stem}.{proj_a}.lora_layer.down.weight"], sd[f"{stem}.{proj_a}.lora_layer.up.weight"] = down, up
        else:            # spelling of LoRAAttnProcessor before its first call
            sd[f"{stem}.processor.{proj_b}_lora.down.weight"], sd[f"{stem}.processor.{proj_b}_lora.up.weight"] = down, up
        expect[name] = w + up @ down
    sd["vae.decoder.conv_in.weight"] = torch.zeros(3)
    sd["eval_metrics.fid.dummy"] = torch.zeros(1)
    return sd, expect


def test_load_reference_state_dict_folds_lora():
    cfg = ou.TINY_CONFIG
    src = synth.build_model(om.MultiViewBaseModel, cfg, seed=0)
    sd, expect = _reference_style(src)
    n_lora = sum(k.endswith("down.weight") for k in sd)
    assert n_lora == 2 * 16 * 8  # 2 UNets x 16 transformer blocks x (attn1, attn2) x (q, k, v, out)
    dst = MultiViewBaseModel(ou.build_unet(cfg, seed=7), ou.build_unet(cfg, seed=8))
    rep = ck.load_reference_state_dict(dst, sd, strict=True)
    assert rep["folded"] == n_lora and not rep["missing"] and not rep["unexpected"]
    got = dst.state_dict()
    assert set(got) == set(expect)
    for name, w in expect.items():
        torch.testing.assert_close(got[name], w, rtol=0, atol=1e-6, msg=name)
    # the folded Linear equals base + LoRA branch on activations (diffusers LoRACompatibleLinear semantics)
    name = "unet.mid_block.attentions.0.transformer_blocks.0.attn1.to_q.weight"
    x = torch.randn(5, got[name].shape[1])
    stem = "mv_base_model.unet._orig_mod.mid_block.attentions.0.transformer_blocks.0.attn1"
    keys = [k for k in sd if k.startswith(stem) and "to_q" in k and "lora" in k]
    down = sd[[k for k in keys if k.endswith("down.weight")][0]]
    up = sd[[k for k in keys if k.endswith("up.weight")][0]]
    torch.testing.assert_close(x @ got[name].T, x @ src.state_dict()[name].T + (x @ down.T) @ up.T, rtol=1e-5, atol=1e-5)


def test_load_reference_state_dict_reports_and_strictness():
    cfg = ou.TINY_CONFIG
    src = synth.build_model(om.MultiViewBaseModel, cfg, seed=0)
    sd, _ = _reference_style(src)
    some = next(k for k in sd if k.endswith("cp_blocks_mid.transformer.norm1.weight"))
    del sd[some]
    sd["mv_base_model.unet._orig_mod.not_a_layer.weight"] = torch.zeros(2)
    dst = MultiViewBaseModel(ou.build_unet(cfg, seed=7), ou.build_unet(cfg, seed=8))
    rep = ck.load_reference_state_dict(dst, sd)
    assert rep["missing"] == ["cp_blocks_mid.transformer.norm1.weight"] and rep["unexpected"] == ["unet.not_a_layer.weight"]
    with pytest.raises(RuntimeError):
        ck.load_reference_state_dict(dst, sd, strict=True)
    bad = dict(sd)
    k = next(k for k in bad if k.endswith("lora_layer.up.weight"))
    del bad[k]
    with pytest.raises(KeyError):
        ck.load_reference_state_dict(dst, bad)


def test_split_handles_both_spellings():
    sd = {"mv_base_model.unet._orig_mod.a.attn1.to_out.0.lora_layer.down.weight": torch.zeros(4, 8),
          "mv_base_model.unet._orig_mod.a.attn1.to_out.0.lora_layer.up.weight": torch.zeros(8, 4),
          "mv_base_model.pano_unet.b.attn2.processor.to_out_lora.down.weight": torch.zeros(4, 8),
          "mv_base_model.pano_unet.b.attn2.processor.to_out_lora.up.weight": torch.zeros(8, 4),
          "mv_base_model.cp_blocks_mid.pe.freq_bands": torch.zeros(3), "text_encoder.x": torch.zeros(1)}
    plain, lora = ck.split_reference_state_dict(sd)
    assert set(plain) == {"cp_blocks_mid.pe.freq_bands"}
    assert set(lora) == {"unet.a.attn1.to_out.0.weight", "pano_unet.b.attn2.to_out.0.weight"}
