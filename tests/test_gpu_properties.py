"""-m gpu: size-independent properties at BASELINE's full sizes (SD-2-size UNets, 64x128 pano + 8 views of 64x64,
CFG batch 2) — where the fp32 CPU oracle would take minutes per step.

  * determinism: two forwards are bit-identical (fixed-order reductions everywhere);
  * batch independence: both CFG halves given the SAME prompt produce identical predictions;
  * identity fusion: with the reference's zero-initialised EPPA output projections (transformer.py:29-30,54-55) the
    fusion blocks are exact identities, so the result cannot depend on the cameras;
  * panorama shift equivariance: the panorama-only model (unet=None, PanoOnly.py:13) with circular padding commutes
    with rolling the latent by a multiple of 8 columns (three stride-2 stages).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sd2(cuda_device):
    from panfusion_b200 import sd2_unet
    return (sd2_unet.build_synthetic(seed=1, device=cuda_device), sd2_unet.build_synthetic(seed=2, device=cuda_device))


def _inputs(dev, m=8, same_prompt=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    theta = torch.tensor(np.rad2deg(np.linspace(0, 2 * np.pi, m, endpoint=False)), dtype=torch.float32)
    cams = dict(FoV=torch.full((2, m), 90.0), theta=theta[None].repeat(2, 1), phi=torch.zeros(2, m))
    pano = torch.randn(1, 1, 4, 64, 128, generator=g).repeat(2, 1, 1, 1, 1)
    lat = torch.randn(1, m, 4, 64, 64, generator=g).repeat(2, 1, 1, 1, 1)
    text, null = torch.randn(1, 1, 77, 1024, generator=g), torch.randn(1, 1, 77, 1024, generator=g)
    if same_prompt:
        null = text
    return dict(latents=lat.to(dev), pano_latent=pano.to(dev), timestep=torch.full((2, m), 501.0, device=dev),
                prompt_embd=torch.cat([null.repeat(1, m, 1, 1), text.repeat(1, m, 1, 1)]).to(dev),
                pano_prompt_embd=torch.cat([null, text]).to(dev), cameras=cams)


def test_determinism_batch_independence_and_identity_fusion(cuda_device, sd2):
    from panfusion_b200.mvgen import MultiViewBaseModel
    torch.manual_seed(0)
    model = MultiViewBaseModel(*sd2, compute_dtype=torch.bfloat16).to(cuda_device).eval()  # EPPA outputs zero-init
    model.prepare(cuda_device, torch.bfloat16)
    inp = _inputs(cuda_device, same_prompt=True)
    s1, p1 = model(**inp)
    s2, p2 = model(**inp)
    assert torch.equal(s1, s2) and torch.equal(p1, p2)                       # determinism
    assert torch.equal(s1[0], s1[1]) and torch.equal(p1[0], p1[1])           # CFG halves independent & identical
    assert torch.isfinite(s1).all() and torch.isfinite(p1).all() and s1.abs().max() > 1e-3
    other = dict(inp)
    other["cameras"] = dict(inp["cameras"], theta=(inp["cameras"]["theta"] + 33.0) % 360,
                            phi=inp["cameras"]["phi"] + 10.0)
    s3, p3 = model(**other)
    assert torch.equal(s1, s3) and torch.equal(p1, p3)                       # zero-init EPPA == identity


def test_pano_only_shift_equivariance(cuda_device, sd2):
    from panfusion_b200.mvgen import MultiViewBaseModel
    model = MultiViewBaseModel(None, sd2[1], compute_dtype=torch.bfloat16).to(cuda_device).eval()
    model.prepare(cuda_device, torch.bfloat16)
    g = torch.Generator().manual_seed(1)
    pano = torch.randn(1, 1, 4, 64, 128, generator=g).to(cuda_device)
    prompt = torch.randn(1, 1, 77, 1024, generator=g).to(cuda_device)
    t = torch.tensor([481.0], device=cuda_device)
    _, a = model(None, pano, t, None, prompt, None)
    _, b = model(None, torch.roll(pano, 40, dims=-1), t, None, prompt, None)
    ref = torch.roll(a, 40, dims=-1)
    err = (b - ref).abs().max().item() / ref.abs().max().item()
    print(f"[property] pano shift equivariance: {err:.3e} of max|out|")
    assert err < 2e-2  # bf16 storage; GroupNorm partial sums are regrouped by the shift
