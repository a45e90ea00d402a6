"""CPU (-m "not gpu"): the C-ABI library builds, loads and exports exactly what include/panfusion_b200.h declares;
host-side logic that needs no GPU (packing, camera records, schedule)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "panfusion_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from panfusion_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.pf_version() >= 100


def test_bad_arguments_fail_loudly_without_gpu():
    """Argument validation happens before any CUDA call: error code + message, mapped to Python exceptions."""
    from panfusion_b200 import _lib
    lib = _lib.lib()
    rc = lib.pf_e2p(None, None, 0, 1, 1, 8, 16, 4, 4, None, 0, 0, None)
    assert rc == -1 and b"null pointer" in lib.pf_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert lib.pf_gemm_pick_block_n(320, 0) == 160 and lib.pf_gemm_pick_block_n(128, 0) == 128
    assert lib.pf_gemm_pick_block_n(100, 0) == 0


def test_no_cpu_path():
    from panfusion_b200 import geometry
    with pytest.raises(Exception):
        geometry.e2p(torch.zeros(1, 1, 8, 16), 90, 0, 0, (4, 4))  # CPU tensor: refused, never computed on the host


def test_camera_record_matches_oracle_rotations():
    from oracle import geometry as og
    from panfusion_b200.geometry import _camera_record
    for fov, th, ph in [(90.0, 0.0, 0.0), (75.0, 45.0, 30.0), (100.0, 200.0, -60.0)]:
        rec = np.array(_camera_record("e2p", fov, th, ph, 16, 24))
        R1, R2 = og.camera_rotations(th, ph)
        np.testing.assert_allclose(rec[:9].reshape(3, 3), R1, atol=1e-15)
        np.testing.assert_allclose(rec[9:18].reshape(3, 3), R2, atol=1e-15)
        assert rec[18] == np.tan(np.radians(fov / 2.0))
        rec = np.array(_camera_record("p2e", fov, th, ph, 16, 24))
        np.testing.assert_allclose(rec[:9].reshape(3, 3), np.linalg.inv(R1), atol=1e-15)


def test_geglu_and_conv_packing():
    from panfusion_b200.packing import pack_conv3x3, pack_geglu
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = pack_conv3x3(w)
    assert p.shape == (2, 27) and p[1, 4 * 3 + 2] == w[1, 2, 1, 1]  # tap (1,1), channel 2
    W = torch.randn(640, 8)
    b = torch.randn(640)
    wp, bp = pack_geglu(W, b, 160)
    assert torch.equal(wp[:80], W[:80]) and torch.equal(wp[80:160], W[320:400]) and torch.equal(wp[160:240], W[80:160])
    assert torch.equal(bp[80:160], b[320:400])


def test_schedule_matches_oracle():
    from oracle.sampler import DDIM
    from panfusion_b200.sampler import DDIMSchedule
    a, b = DDIMSchedule(), DDIM()
    a.set_timesteps(50)
    b.set_timesteps(50)
    assert torch.equal(a.timesteps, b.timesteps)
    x, e = torch.randn(8, dtype=torch.float64), torch.randn(8, dtype=torch.float64)
    for t in (981, 501, 1):
        cx, ce = a.coefficients(t)
        torch.testing.assert_close(cx * x + ce * e, b.step(e, t, x).double(), rtol=1e-5, atol=1e-6)


def test_camera_table_dedup():
    from panfusion_b200.eppa import CameraTables
    cams = dict(FoV=torch.full((4,), 90.0), theta=torch.tensor([0.0, 180.0, 0.0, 180.0]), phi=torch.zeros(4))
    key, groups = CameraTables.dedup(CameraTables.camera_key(cams), 2)
    assert groups == 1 and len(key[0]) == 2
    cams["theta"] = torch.tensor([0.0, 180.0, 90.0, 270.0])
    key, groups = CameraTables.dedup(CameraTables.camera_key(cams), 2)
    assert groups == 2 and len(key[0]) == 4


def test_row_stats_slots_ignore_tile_requests():
    """A fused-LayerNorm PRODUCER's slot layout is a function of N alone (never of a tile-width request or of tuning), and the
    query gives the same answer before and after the caller has filled in row_stats_out."""
    import ctypes as C
    from panfusion_b200 import _lib
    lib = _lib.lib()
    for n, want in ((320, 4), (640, 8), (1280, 16), (128, 2), (64, 2)):
        seen = set()
        for req in (0, 64, 128, 64 | (2 << 16), 256 | (1 << 16)):
            a = _lib.GemmArgs()
            a.N, a.M, a.Kc, a.num_taps, a.block_n = n, 512, n, 1, req
            seen.add(lib.pf_gemm_row_stats_slots(C.byref(a)))
            a.row_stats_out = 1
            seen.add(lib.pf_gemm_row_stats_slots(C.byref(a)))
        assert seen == {want}, (n, seen)


def test_packed_bias_tile_layout_helper():
    """ops.bias_tile_dense inverts the lane-interleaved tile layout documented at pf_bias_tile_pack
    (include/panfusion_b200.h): element (r, c) of a 128 x 64 tile sits at (((r/32)*16 + c/4)*32 + r%32)*4 + c%4."""
    import torch
    from panfusion_b200 import ops
    dense = torch.arange(128 * 64, dtype=torch.float32).reshape(128, 64)
    r, c = torch.meshgrid(torch.arange(128), torch.arange(64), indexing="ij")
    pos = (((r // 32) * 16 + c // 4) * 32 + r % 32) * 4 + c % 4
    assert sorted(pos.flatten().tolist()) == list(range(128 * 64))  # a permutation of the tile
    stored = torch.empty(128 * 64)
    stored[pos.flatten()] = dense.flatten()
    assert torch.equal(ops.bias_tile_dense(stored), dense)
    # one 16-byte piece of a warp = 32 lanes x 4 floats, contiguous
    w, e = 2, 5
    piece = stored[((w * 16 + e) * 32) * 4:((w * 16 + e) * 32 + 32) * 4].reshape(32, 4)
    assert torch.equal(piece, dense[w * 32:(w + 1) * 32, e * 4:(e + 1) * 4])
