"""CPU (-m "not gpu"): the oracle restatement against the goldens minted by executing the reference's own files
(oracle/make_golden.py, run in the build container) — this is what pins the oracle."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import eppa as oe, geometry as og, mvgen as om, synth, unet as ou

GOLD = Path(__file__).parent / "golden"


def _cams3():
    return dict(FoV=torch.tensor([90.0, 75.0, 100.0]), theta=torch.tensor([0.0, 45.0, 200.0]),
                phi=torch.tensor([0.0, 30.0, -60.0]))


def test_resample_matches_reference_golden():
    gold = np.load(GOLD / "resample.npz")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 5, 32, 64, generator=g)
    y = torch.randn(3, 5, 16, 24, generator=g)
    c = _cams3()
    for mode in ("bilinear", "nearest"):
        out = og.e2p(x, c["FoV"], c["theta"], c["phi"], (16, 24), mode=mode)
        np.testing.assert_array_equal(out.numpy(), gold[f"e2p_{mode}"])
        out, mask = og.p2e(y, c["FoV"], c["theta"], c["phi"], (32, 64), mode=mode)
        np.testing.assert_array_equal(out.numpy(), gold[f"p2e_{mode}"])
        np.testing.assert_array_equal(mask.numpy(), gold[f"p2e_{mode}_mask"])
    np.testing.assert_array_equal(og.e2p(x, 90, 10, 5, (16, 16)).numpy(), gold["e2p_scalar"])


def test_eppa_geometry_matches_reference_golden():
    gold = np.load(GOLD / "eppa_geometry.npz")
    c = _cams3()
    pm, em = oe.get_masks(8, 8, 8, 16, c)
    pc, ec = oe.get_coords(8, 8, 8, 16, c)
    np.testing.assert_allclose(pm.numpy(), gold["pers_masks"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(em.numpy(), gold["equi_masks"], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(pc.numpy(), gold["pers_coords"])
    np.testing.assert_array_equal(ec.numpy(), gold["equi_coords"])
    # structure the kernels rely on: rows are -1 except near correspondences, row max is exactly +1
    assert pm.min() >= -1 and pm.max() <= 1 and em.max() == 1.0


def test_eppa_geometry_config4_matches_reference_golden():
    """ph != eh and icosahedron-ring cameras (BASELINE config 4's rig): golden from the reference's get_masks/get_coords."""
    gold = np.load(GOLD / "eppa_geometry_c4.npz")
    c = dict(FoV=torch.full((4,), 90.0), theta=torch.tensor([-144.0, 72.0, -180.0, 36.0]),
             phi=torch.tensor([52.6226, 10.8123, -10.8123, -52.6226]))
    pm, em = oe.get_masks(8, 8, 16, 32, c)
    pc, ec = oe.get_coords(8, 8, 16, 32, c)
    np.testing.assert_allclose(pm.numpy(), gold["pers_masks"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(em.numpy(), gold["equi_masks"], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(pc.numpy(), gold["pers_coords"])
    np.testing.assert_array_equal(ec.numpy(), gold["equi_coords"])


def test_eppa_geometry_config4_level_size_matches_reference_golden():
    """get_masks at config 4's REAL first EPPA level (32x32 views / 64x128 pano, one camera per icosahedron ring):
    the fixture holds a strided subset of the query rows plus the key-sum of EVERY row of the reference output."""
    from oracle.make_golden import C4GEO_LEVEL, _cams_ico, c4geo_subsample
    gold = np.load(GOLD / "eppa_geometry_c4_level.npz")
    pm, em = oe.get_masks(*C4GEO_LEVEL, _cams_ico())
    mine = c4geo_subsample(pm, em)
    for k in ("pers_rows", "equi_rows"):
        np.testing.assert_allclose(mine[k], gold[k], rtol=0, atol=1e-6)
    for k in ("pers_rowsum", "equi_rowsum"):
        np.testing.assert_allclose(mine[k], gold[k], rtol=0, atol=1e-3)


def test_mask_edge_cases():
    """Camera looking at the pole / FoV so narrow that many queries have no correspondence: rows stay -1."""
    c = dict(FoV=torch.tensor([30.0]), theta=torch.tensor([10.0]), phi=torch.tensor([85.0]))
    pm, em = oe.get_masks(4, 4, 8, 16, c)
    rows = pm.reshape(8 * 16, -1)
    empty = (rows.max(dim=1).values == -1)
    assert empty.any() and (~empty).any()
    assert torch.all(rows[~empty].max(dim=1).values == 1.0)


def test_warpattn_matches_reference_golden():
    gold = np.load(GOLD / "warpattn_320.npz")
    torch.manual_seed(7)
    w = oe.WarpAttn(320).eval()
    holder = torch.nn.Module()
    holder.cp_blocks = w
    synth.randomize_zero_init(holder, 11)
    g = torch.Generator().manual_seed(8)
    px, ex = torch.randn(4, 320, 8, 8, generator=g), torch.randn(2, 320, 8, 16, generator=g)
    c4 = dict(FoV=torch.full((4,), 90.0), theta=torch.tensor([0.0, 180.0, 0.0, 180.0]), phi=torch.zeros(4))
    with torch.no_grad():
        p, e = w(px, ex, c4)
    np.testing.assert_allclose(p.numpy(), gold["pers_out"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(e.numpy(), gold["equi_out"], rtol=1e-5, atol=1e-5)
    assert (p - px).abs().max() > 0.1  # the redrawn zero-init tensors make the block a non-identity


def test_mvgen_tiny_matches_reference_golden():
    gold = np.load(GOLD / "mvgen_tiny.npz")
    model = synth.build_model(om.MultiViewBaseModel, ou.TINY_CONFIG, seed=0)
    inp = synth.step_inputs(2, (16, 32), (16, 16), ou.TINY_CONFIG["cross_attention_dim"], seed=0)
    with torch.no_grad():
        s, p = model(**inp)
    np.testing.assert_allclose(s.numpy(), gold["sample"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(p.numpy(), gold["pano_sample"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("tag,pers", [("tiny_cn", False), ("tiny_cn2", True)])
def test_mvgen_controlnet_matches_reference_golden(tag, pers):
    """Layout-conditioned step (BASELINE config 5): golden = reference MVGenModel.py run around the ControlNet
    restatement; the condition must change the output (zero convs are redrawn)."""
    gold = np.load(GOLD / f"mvgen_{tag}.npz")
    model = synth.build_model_cn(om.MultiViewBaseModel, ou.TINY_CONFIG, seed=0, pers=pers)
    inp = synth.step_inputs(2, (16, 32), (16, 16), ou.TINY_CONFIG["cross_attention_dim"], seed=0)
    inp.update(synth.layout_conds(1, 2, (16, 32), (16, 16), seed=5, pers=pers))
    with torch.no_grad():
        s, p = model(**inp)
        s0, p0 = model(**{**inp, "pano_layout_cond": None, "pers_layout_cond": None})
    np.testing.assert_allclose(s.numpy(), gold["sample"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(p.numpy(), gold["pano_sample"], rtol=1e-4, atol=1e-4)
    assert (p - p0).abs().max() > 0.05
    base = np.load(GOLD / "mvgen_tiny.npz")
    np.testing.assert_allclose(p0.numpy(), base["pano_sample"], rtol=1e-4, atol=1e-4)  # cond=None == no ControlNet


def test_pano_only_branch():
    """unet=None (PanoOnly ablation, models/pano/PanoOnly.py:13): timestep is [b], no EPPA blocks."""
    pano_unet = ou.build_unet(ou.TINY_CONFIG, seed=2)
    model = om.MultiViewBaseModel(None, pano_unet).eval()
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        s, p = model(None, torch.randn(1, 1, 4, 16, 32, generator=g), torch.tensor([981]), None,
                     torch.randn(1, 1, 77, 64, generator=g), None)
    assert s is None and p.shape == (1, 1, 4, 16, 32)


def test_pad_pano_errors_and_roundtrip():
    x = torch.randn(2, 3, 4, 8)
    assert torch.equal(oe.unpad_pano(oe.pad_pano(x, 2), 2), x)
    assert torch.equal(oe.pad_pano(x, 2)[..., :2], x[..., -2:])
    assert oe.pad_pano(x, 0) is x
    with pytest.raises(NotImplementedError):
        oe.pad_pano(torch.randn(4, 8), 1)


def test_ddim_schedule_matches_sd2_leading_spacing():
    from oracle.sampler import DDIM
    s = DDIM()
    s.set_timesteps(50)
    assert s.timesteps[0] == 981 and s.timesteps[-1] == 1 and len(s.timesteps) == 50
    x, e = torch.randn(4), torch.randn(4)
    a_t, a_p = s.alphas_cumprod[981], s.alphas_cumprod[961]
    ref = a_p.sqrt() * (x - (1 - a_t).sqrt() * e) / a_t.sqrt() + (1 - a_p).sqrt() * e
    torch.testing.assert_close(s.step(e, 981, x), ref)


def test_oracle_vae_tail_first_party_semantics():
    """decode_latent / padded panorama decode / tensor_to_image (PanoGenerator.py:272-278, PanFusion.py:166-172,
    models/modules/utils.py:9-15) around the [3P] decoder restatement: shapes, scaling, circular seam, rounding."""
    from oracle import vae as ov
    vae = ov.build_vae(ov.TINY_VAE_CONFIG)
    g = torch.Generator().manual_seed(0)
    pano = torch.randn(1, 1, 4, 8, 16, generator=g)
    with torch.no_grad():
        img = ov.decode_pano(pano, vae, 8)
        assert img.shape == (1, 1, 3, 64, 128)
        # scaling: decode_latent(z) == vae.decode(z / scaling_factor)
        direct = vae.decode(pano[:, 0] / vae.config.scaling_factor).sample
        torch.testing.assert_close(ov.decode_latent(pano, vae)[:, 0], direct)
        # circular padding makes the decode commute with a roll of the latent up to the padded context
        rolled = ov.decode_pano(torch.roll(pano, 4, dims=-1), vae, 8)
        assert (torch.roll(img, 32, dims=-1) - rolled).abs().max() < 0.35 * img.abs().max()
    x = torch.tensor([[[[-1.0, 1.0, 0.0, -7.0, 7.0, 1 / 255]]]])  # 0.0 -> 127.5 -> 128 (half-to-even, torch.round)
    assert ov.tensor_to_image(x).tolist() == [[[[0], [255], [128], [0], [255], [128]]]]
    full = ov.build_vae()
    n_dec = sum(p.numel() for m in (full.decoder, full.post_quant_conv) for p in m.parameters())
    assert n_dec == 49490199                                             # the SD VAE decoder (+ post_quant_conv)
    assert sum(p.numel() for p in full.parameters()) == 83653863         # the whole SD AutoencoderKL (published size)
    # encode_image / padded panorama encode (PanoGenerator.py:214-225, PanFusion.py:66-71): shapes, scaling, the draw
    g = torch.Generator().manual_seed(1)
    imgs = torch.rand(1, 2, 3, 64, 64, generator=g) * 2 - 1
    noise = torch.randn(2, 4, 8, 8, generator=g)
    with torch.no_grad():
        dist = vae.encode(imgs[0]).latent_dist
        z = ov.encode_image(imgs, vae, noise=noise)
        assert z.shape == (1, 2, 4, 8, 8)
        torch.testing.assert_close(z[0], (dist.mean + dist.std * noise) * vae.config.scaling_factor)
        zp = ov.encode_pano(torch.rand(1, 1, 3, 64, 128, generator=g) * 2 - 1, vae, 8, generator=g)
        assert zp.shape == (1, 1, 4, 8, 16)


def test_py360convert_e2p_matches_reference_golden_and_scipy():
    """Dataset-path convention (external/py360convert/e2p.py:6-43): the restatement against the golden minted from the
    reference file, and its hand-written 'wrap' interpolation against scipy.ndimage.map_coordinates itself."""
    from scipy.ndimage import map_coordinates
    from oracle import py360 as op
    from oracle.make_golden import PY360_CASES, py360_images
    gold = np.load(GOLD / "py360_e2p.npz")
    for k, (fov, u, v, hw, rot, mode) in enumerate(PY360_CASES):
        for tag, im in zip(("u8", "f32"), py360_images()):
            np.testing.assert_array_equal(op.e2p(im, fov, u, v, hw, rot, mode), gold[f"case{k}_{tag}"])
    rng = np.random.default_rng(3)
    img = rng.random((9, 13))
    cy, cx = rng.uniform(-20, 30, 5000), rng.uniform(-30, 40, 5000)
    cy[:40], cx[:40] = rng.integers(-9, 18, 40), rng.integers(-13, 26, 40)
    padded = np.concatenate([img, np.roll(img[[-1]], 13 // 2, 1), np.roll(img[[0]], 13 // 2, 1)], 0)
    for order in (0, 1):
        ref = map_coordinates(padded, [cy, cx], order=order, mode="wrap")
        np.testing.assert_array_equal(op.sample_equirec(img, cx, cy, order), ref)
    with pytest.raises(NotImplementedError):
        op.e2p(img, (90, 90), 0, 0, (4, 4), mode="bicubic")


def test_text_encoder_oracle_is_transformers_clip_and_causal():
    """The text-encoder oracle EXECUTES transformers.CLIPTextModel (the reference's class, PanoGenerator.py:117-121) with the
    SD-2 text-tower shape; check what the GPU tests rely on: seeded determinism, causality, final LayerNorm applied."""
    from oracle import text_encoder as ot
    import transformers
    m = ot.build_text_encoder(ot.TINY_TEXT_CONFIG, seed=0)
    assert isinstance(m, transformers.CLIPTextModel)
    ids = ot.token_ids(2, vocab=1000, seed=1)
    a = ot.encode_text(m, ids)
    b = ot.encode_text(ot.build_text_encoder(ot.TINY_TEXT_CONFIG, seed=0), ids)
    assert torch.equal(a, b) and a.shape == (2, 77, 128)
    ids2 = ids.clone()
    ids2[0, 30] = (ids2[0, 30] + 1) % 998
    c = ot.encode_text(m, ids2)
    assert torch.equal(c[0, :30], a[0, :30]) and not torch.equal(c[0, 30:], a[0, 30:])
    assert ot.SD2_TEXT_CONFIG["num_hidden_layers"] == 23 and ot.SD2_TEXT_CONFIG["hidden_size"] // ot.SD2_TEXT_CONFIG["num_attention_heads"] == 64
