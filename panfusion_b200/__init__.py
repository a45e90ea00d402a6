"""panfusion_b200 — B200-native (sm_100a) denoise hot path of PanFusion behind the reference's interfaces."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
