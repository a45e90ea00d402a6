"""Load the reference's OWN hot-path files by path (build container only; `/root/reference` is absent on the
GPU box). Used by oracle/make_golden.py to pin the restatement and mint tests/golden/*.

`models/__init__.py` eagerly imports lightning/diffusers (not installed), so empty package objects are
registered for `models`, `models.pano`, `models.modules`; kornia / xformers resolve to oracle/shims.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

REF = Path("/root/reference")
SHIMS = Path(__file__).resolve().parent / "shims"


def available() -> bool:
    return (REF / "models" / "pano" / "MVGenModel.py").exists()


def load():
    """-> namespace with e2p, p2e, map_pers_coords_to_equi, get_masks, get_coords, pad_pano, unpad_pano,
    transformer (module), WarpAttn, MultiViewBaseModel from the reference."""
    if not available():
        raise RuntimeError("/root/reference is not mounted (GPU box?) - the reference loader only works in "
                           "the build container")
    for p in (str(REF), str(SHIMS)):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if "external.PanoAnnotator" not in sys.modules:
        sys.modules["external.PanoAnnotator"] = types.ModuleType("external.PanoAnnotator")
    for name, sub in (("models", "models"), ("models.pano", "models/pano"), ("models.modules", "models/modules")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [str(REF / sub)]
            sys.modules[name] = m
    ns = types.SimpleNamespace()
    pe = importlib.import_module("external.Perspective_and_Equirectangular")
    ns.e2p, ns.p2e, ns.map_pers_coords_to_equi = pe.e2p, pe.p2e, pe.map_pers_coords_to_equi
    up = importlib.import_module("utils.pano")
    ns.pad_pano, ns.unpad_pano = up.pad_pano, up.unpad_pano
    ns.horizon_sample_camera, ns.icosahedron_sample_camera = up.horizon_sample_camera, up.icosahedron_sample_camera
    ns.transformer = importlib.import_module("models.modules.transformer")
    pu = importlib.import_module("models.pano.utils")
    ns.get_masks, ns.get_coords = pu.get_masks, pu.get_coords
    ns.WarpAttn = importlib.import_module("models.pano.modules").WarpAttn
    ns.MultiViewBaseModel = importlib.import_module("models.pano.MVGenModel").MultiViewBaseModel
    return ns
