"""-m gpu: the drop-in modules (WarpAttn, MultiViewBaseModel) through the full CUDA path against
(a) the committed goldens minted by executing the reference's own files (oracle/make_golden.py) and
(b) the CPU oracle on the same seeded inputs.

Tolerances. The north star asks rtol 1e-3 / atol 1e-4 "fp16"; that bar is met per kernel (tests/test_gpu_gemm.py,
test_gpu_fmha.py, test_gpu_kernels.py, test_gpu_resample.py compare each kernel with an fp32 reference on
16-bit-rounded inputs). End to end the activations are ROUNDED TO 16 BIT between ~400 kernels, which the fp32
reference never does, so the whole-model comparison is bounded by accumulated storage rounding instead:
fp16 (11-bit significand)  : max |err| <= 1.5e-2 * max|ref|, mean |err| <= 2e-3 * max|ref|
bf16 ( 8-bit significand)  : max |err| <= 8e-2  * max|ref|, mean |err| <= 1.2e-2 * max|ref|
(measured values are printed; see DESIGN.md "Parity").
"""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _err(got, ref):
    scale = ref.abs().max().item()
    d = (got - ref).abs()
    return d.max().item() / scale, d.mean().item() / scale


def _check(name, got, ref, dtype):
    mx, mean = _err(got.float().cpu(), ref)
    lim = (1.5e-2, 2e-3) if dtype == torch.float16 else (8e-2, 1.2e-2)
    print(f"[parity] {name} {dtype}: max {mx:.3e} mean {mean:.3e} (of max|ref|) limits {lim}")
    assert mx <= lim[0] and mean <= lim[1], (name, mx, mean)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_warpattn_vs_reference_golden(cuda_device, dtype):
    """WarpAttn(320) on 2 batches x 2 views, 8x8 / 8x16 — golden = reference modules.py:8-59 output."""
    from oracle import eppa as oe, synth
    from panfusion_b200.eppa import WarpAttn
    torch.manual_seed(7)
    worc = oe.WarpAttn(320).eval()
    holder = torch.nn.Module()
    holder.cp_blocks = worc
    synth.randomize_zero_init(holder, 11)
    mine = WarpAttn(320).eval()
    mine.load_state_dict(worc.state_dict())
    g = torch.Generator().manual_seed(8)
    px, ex = torch.randn(4, 320, 8, 8, generator=g), torch.randn(2, 320, 8, 16, generator=g)
    c4 = dict(FoV=torch.full((4,), 90.0), theta=torch.tensor([0.0, 180.0, 0.0, 180.0]), phi=torch.zeros(4))
    gold = np.load(GOLD / "warpattn_320.npz")
    with torch.no_grad():
        op, oq = worc(px, ex, c4)
    # the oracle reproduces the reference golden (pins the oracle on this box too)
    torch.testing.assert_close(op, torch.from_numpy(gold["pers_out"]), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(oq, torch.from_numpy(gold["equi_out"]), rtol=1e-5, atol=1e-5)
    gp, ge = mine.to(cuda_device)(px.to(cuda_device), ex.to(cuda_device), c4, compute_dtype=dtype)
    # the block is residual: compare the UPDATE it adds, which is what the kernels compute
    _check("WarpAttn pers update", gp.float().cpu() - px, torch.from_numpy(gold["pers_out"]) - px, dtype)
    _check("WarpAttn equi update", ge.float().cpu() - ex, torch.from_numpy(gold["equi_out"]) - ex, dtype)


def _run_mvgen(cuda_device, config, pano_hw, pers_hw, dtype):
    from oracle import mvgen as om, synth
    from panfusion_b200.mvgen import MultiViewBaseModel
    orc = synth.build_model(om.MultiViewBaseModel, config, seed=0)
    inp = synth.step_inputs(2, pano_hw, pers_hw, config["cross_attention_dim"], seed=0)
    mine = MultiViewBaseModel(orc.unet, orc.pano_unet, compute_dtype=dtype)
    mine.load_state_dict(orc.state_dict())
    mine.prepare(cuda_device, dtype)
    cu = {k: (v.to(cuda_device) if torch.is_tensor(v) else {kk: vv.to(cuda_device) for kk, vv in v.items()})
          for k, v in inp.items()}
    s, p = mine(**cu)
    torch.cuda.synchronize()
    return orc, inp, s, p


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_mvgen_tiny_vs_golden_and_oracle(cuda_device, dtype):
    from oracle import unet as ou
    orc, inp, s, p = _run_mvgen(cuda_device, ou.TINY_CONFIG, (16, 32), (16, 16), dtype)
    gold = np.load(GOLD / "mvgen_tiny.npz")
    with torch.no_grad():
        os_, op_ = orc(**inp)
    torch.testing.assert_close(os_, torch.from_numpy(gold["sample"]), rtol=1e-4, atol=1e-4)
    assert s.shape == os_.shape and p.shape == op_.shape
    _check("MultiViewBaseModel tiny sample", s, torch.from_numpy(gold["sample"]), dtype)
    _check("MultiViewBaseModel tiny pano", p, torch.from_numpy(gold["pano_sample"]), dtype)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_mvgen_c1_vs_reference_golden(cuda_device, dtype):
    """BASELINE config 1 (SD-2-size UNets, 1 pano 64x128 + 2 views 64x64): golden = the reference's own
    MultiViewBaseModel.forward executed on CPU in the build container."""
    from oracle import unet as ou
    _, _, s, p = _run_mvgen(cuda_device, ou.SD2_CONFIG, (64, 128), (64, 64), dtype)
    gold = np.load(GOLD / "mvgen_c1.npz")
    _check("MultiViewBaseModel C1 sample", s, torch.from_numpy(gold["sample"]), dtype)
    _check("MultiViewBaseModel C1 pano", p, torch.from_numpy(gold["pano_sample"]), dtype)
