"""kornia.filters.gaussian_blur2d [3P restatement] — call sites models/pano/utils.py:65,67.
kornia 0.7.2 default separable=True: pad (border_type) then two 1-D depthwise convolutions."""
import torch
import torch.nn.functional as F


def _kernel1d(ksize, sigma, device, dtype):
    x = torch.arange(ksize, device=device, dtype=dtype) - ksize // 2
    if ksize % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
    return g / g.sum()


def gaussian_blur2d(input, kernel_size, sigma, border_type="reflect", separable=True):
    ky, kx = kernel_size
    sy, sx = float(sigma[0]), float(sigma[1])
    b, c, h, w = input.shape
    k_x = _kernel1d(kx, sx, input.device, input.dtype)
    k_y = _kernel1d(ky, sy, input.device, input.dtype)
    # filter2d_separable: filter2d(x, kernel_x[None]) then filter2d(., kernel_y[..., None]); each pads its own axis
    out = F.pad(input, [kx // 2, kx // 2, 0, 0], mode=border_type)
    out = F.conv2d(out.reshape(-1, 1, h, w + 2 * (kx // 2)), k_x.view(1, 1, 1, kx)).reshape(b, c, h, w)
    out = F.pad(out, [0, 0, ky // 2, ky // 2], mode=border_type)
    out = F.conv2d(out.reshape(-1, 1, h + 2 * (ky // 2), w), k_y.view(1, 1, ky, 1)).reshape(b, c, h, w)
    return out
