"""kornia.geometry.transform.remap [3P restatement] — call sites e2p.py:76, p2e.py:70."""
import torch
import torch.nn.functional as F


def remap(image, map_x, map_y, mode="bilinear", padding_mode="zeros", align_corners=None,
          normalized_coordinates=False):
    b, _, h, w = image.shape
    grid = torch.stack([map_x, map_y], dim=-1)
    if not normalized_coordinates:
        hw = torch.tensor([w, h], device=grid.device, dtype=grid.dtype)
        factor = torch.tensor(2.0, device=grid.device, dtype=grid.dtype) / (hw - 1).clamp(1e-14)
        grid = factor * grid - 1
    grid = grid.expand(b, -1, -1, -1)
    return F.grid_sample(image, grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)
