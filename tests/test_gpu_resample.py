"""-m gpu: spherical resampling kernels (pf_e2p / pf_p2e) against the CPU oracle (oracle/geometry.py).

fp32: the kernel rounds the fp64 grid to fp32 at the same point as the reference and replays kornia's /
ATen's fp32 normalise-unnormalise, so results agree to fp32 round-off: rtol 1e-5 / atol 1e-6 (well inside the
north-star rtol 1e-3 / atol 1e-4). nearest mode must select the same source pixel (exact equality, except
ties within 1 ulp of .5 which the tolerance on mismatch count covers: none observed).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # B, C, He, We, h, w
    (2, 4, 64, 128, 64, 64),      # init_noise (PanFusion.py:30-43)
    (3, 37, 32, 64, 32, 32),      # staged path, ragged channel count
    (2, 2048, 32, 64, 32, 32),    # get_masks e2p shape (models/pano/utils.py:38-41), staged
    (1, 3, 512, 1024, 256, 256),  # pixel-space panorama: direct path
    (2, 5, 16, 32, 8, 8),
    (1, 1, 8, 16, 4, 6),          # non-square perspective
]


def _cams(B):
    rng = np.random.default_rng(5)
    fov = torch.tensor(rng.uniform(60, 110, B), dtype=torch.float32)
    theta = torch.tensor(rng.uniform(-180, 360, B), dtype=torch.float32)
    phi = torch.tensor(rng.uniform(-80, 80, B), dtype=torch.float32)
    return fov, theta, phi


@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
@pytest.mark.parametrize("B,C,He,We,h,w", CASES)
def test_e2p_fp32(cuda_device, mode, B, C, He, We, h, w):
    from oracle import geometry as og
    from panfusion_b200 import geometry as pg
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, C, He, We, generator=g)
    fov, theta, phi = _cams(B)
    ref = og.e2p(x, fov, theta, phi, (h, w), mode=mode)
    got = pg.e2p(x.to(cuda_device), fov, theta, phi, (h, w), mode=mode).cpu()
    if mode == "nearest":
        assert (got != ref).float().mean().item() < 1e-4
    else:
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
@pytest.mark.parametrize("B,C,He,We,h,w", CASES)
def test_p2e_fp32(cuda_device, mode, B, C, He, We, h, w):
    from oracle import geometry as og
    from panfusion_b200 import geometry as pg
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, C, h, w, generator=g)
    fov, theta, phi = _cams(B)
    ref, rmask = og.p2e(x, fov, theta, phi, (He, We), mode=mode)
    got, gmask = pg.p2e(x.to(cuda_device), fov, theta, phi, (He, We), mode=mode)
    assert gmask.dtype == torch.bool and gmask.shape == rmask.shape
    assert (gmask.cpu() != rmask).sum().item() == 0
    if mode == "nearest":
        assert (got.cpu() != ref).float().mean().item() < 1e-4
    else:
        torch.testing.assert_close(got.cpu(), ref, rtol=1e-5, atol=1e-6)


def test_scalar_camera_broadcast(cuda_device):
    """All-scalar camera arguments share one grid over the batch (e2p.py:65-66, p2e.py:56-57)."""
    from oracle import geometry as og
    from panfusion_b200 import geometry as pg
    x = torch.randn(4, 3, 32, 64, generator=torch.Generator().manual_seed(3))
    ref = og.e2p(x, 90, 45, 10, (16, 16))
    got = pg.e2p(x.to(cuda_device), 90, 45, 10, (16, 16)).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    y = torch.randn(4, 3, 16, 16, generator=torch.Generator().manual_seed(4))
    ref, rm = og.p2e(y, 90, 45, 10, (32, 64))
    got, gm = pg.p2e(y.to(cuda_device), 90, 45, 10, (32, 64))
    assert gm.shape == rm.shape == (1, 1, 32, 64)
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_e2p_half(cuda_device, dtype):
    """16-bit feature maps: grid stays fp32 (documented deviation: the reference would round the grid to the
    image dtype, e2p.py:74-75); compare with the fp32 oracle on the rounded input, one output ulp."""
    from oracle import geometry as og
    from panfusion_b200 import geometry as pg
    x = torch.randn(2, 320, 64, 128, generator=torch.Generator().manual_seed(5)).to(dtype)
    fov, theta, phi = _cams(2)
    ref = og.e2p(x.float(), fov, theta, phi, (64, 64))
    got = pg.e2p(x.to(cuda_device), fov, theta, phi, (64, 64)).float().cpu()
    tol = dict(rtol=2 ** -8, atol=2 ** -8) if dtype == torch.bfloat16 else dict(rtol=2 ** -11, atol=2 ** -11)
    torch.testing.assert_close(got, ref, **tol)


def test_bad_mode(cuda_device):
    from panfusion_b200 import geometry as pg
    with pytest.raises(ValueError):
        pg.e2p(torch.zeros(1, 1, 8, 16, device=cuda_device), 90, 0, 0, (4, 4), mode="bicubic")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_e2p_views_per_image_equals_expanded_source(cuda_device, dtype):
    """pf_e2p_shared: m cameras per panorama == e2p of the panorama expanded to its m views (PanFusion.py:33-37),
    bit for bit, for the staged quad path (feature-map shape) and a tiled large output plane."""
    from panfusion_b200 import geometry as pg
    for (bs, m, C, He, We, h, w) in ((2, 8, 320, 64, 128, 64, 64), (1, 3, 5, 32, 64, 96, 96)):
        x = torch.randn(bs, C, He, We, generator=torch.Generator().manual_seed(6)).to(dtype).to(cuda_device)
        fov, theta, phi = _cams(bs * m)
        for mode in ("bilinear", "nearest"):
            a = pg.e2p(x, fov, theta, phi, (h, w), mode=mode, views_per_image=m)
            b = pg.e2p(x.repeat_interleave(m, 0), fov, theta, phi, (h, w), mode=mode)
            assert a.shape == (bs * m, C, h, w) and torch.equal(a, b)
    with pytest.raises(ValueError):
        pg.e2p(x, 90, 0, 0, (8, 8), views_per_image=3)


def test_py360convert_e2p_vs_reference_golden(cuda_device):
    """pf_e2p_py360 (dataset-path convention, external/py360convert/e2p.py:6-43) against the golden minted by executing the
    reference file: uint8 within one level on < 1e-4 of the pixels (float64 atan2 / tan of the device vs. the host libm
    can move a value across a rounding boundary), fp32 images to 1e-6; `nearest` picks the same source pixel."""
    from pathlib import Path
    from oracle.make_golden import PY360_CASES, py360_images
    from panfusion_b200 import py360
    gold = np.load(Path(__file__).parent / "golden" / "py360_e2p.npz")
    for k, (fov, u, v, hw, rot, mode) in enumerate(PY360_CASES):
        for tag, im in zip(("u8", "f32"), py360_images()):
            ref = gold[f"case{k}_{tag}"]
            got = py360.e2p(im, fov, u, v, hw, in_rot_deg=rot, mode=mode)                 # numpy in -> numpy out
            assert isinstance(got, np.ndarray) and got.shape == ref.shape and got.dtype == ref.dtype
            d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
            if tag == "u8" or mode == "nearest":
                lim = 1 if mode == "bilinear" else 255
                assert d.max() <= lim and (d > 0).mean() < 1e-4, (k, tag, d.max(), (d > 0).mean())
            else:
                assert d.max() < 1e-6, (k, tag, d.max())
    # all views of a panorama in one launch, CUDA tensor in -> CUDA tensor out, 2-D image
    im = py360_images()[0]
    t = torch.from_numpy(im).to(cuda_device)
    yaws, pitches = [0.0, 45.0, 200.0, -170.0], [0.0, 30.0, -60.0, 85.0]
    multi = py360.e2p_views(t, (90, 90), yaws, pitches, (32, 32))
    assert multi.is_cuda and multi.shape == (4, 32, 32, 3)
    for i in range(4):
        assert torch.equal(multi[i], py360.e2p(t, (90, 90), yaws[i], pitches[i], (32, 32)))
    g2 = py360.e2p(im[..., 0], (90, 90), 30.0, 20.0, (48, 48))
    assert g2.shape == (48, 48) and np.abs(g2.astype(int) - gold["case0_u8"][..., 0].astype(int)).max() <= 1
    with pytest.raises(NotImplementedError):
        py360.e2p(im, (90, 90), 0, 0, (4, 4), mode="bicubic")
