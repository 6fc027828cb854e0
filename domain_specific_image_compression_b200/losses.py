"""Distortion terms of rate_distortion_loss (R2 in SURVEY.md 8(a)); these stay in PyTorch (3-channel images, small).

`multi_scale_ssim` restates piq 0.8.0's `multi_scale_ssim` (called at /root/reference/code/modelv2/model.py:96-101 with
data_range=1, scale_weights=[.3,.5,.2]); piq is a third-party package that is neither vendored with the reference nor
installed here, so this follows its published algorithm: Gaussian 11x11 window (sigma 1.5), k1=.01, k2=.03, "valid"
depthwise convolutions, 2x2 average pooling between scales (replicate-padding odd sizes on the top/left), relu on the
per-scale terms, prod_i cs_i^w_i * ssim_last^w_last, mean over channels then batch.  PARITY UNPINNED.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _gaussian_window(size: int, sigma: float, device, dtype) -> torch.Tensor:
    coords = torch.arange(size, dtype=dtype, device=device) - (size - 1) / 2.0
    g = torch.exp(-(coords ** 2) / (2.0 * sigma ** 2))
    g = g / g.sum()
    return (g[:, None] * g[None, :])[None, None]          # [1,1,k,k], sums to 1


def _ssim_and_cs(x, y, win, c1, c2):
    C = x.size(1)
    mu_x = F.conv2d(x, win, groups=C)
    mu_y = F.conv2d(y, win, groups=C)
    mu_xx, mu_yy, mu_xy = mu_x * mu_x, mu_y * mu_y, mu_x * mu_y
    s_xx = F.conv2d(x * x, win, groups=C) - mu_xx
    s_yy = F.conv2d(y * y, win, groups=C) - mu_yy
    s_xy = F.conv2d(x * y, win, groups=C) - mu_xy
    cs = (2.0 * s_xy + c2) / (s_xx + s_yy + c2)
    ss = (2.0 * mu_xy + c1) / (mu_xx + mu_yy + c1) * cs
    return ss.mean(dim=(-1, -2)), cs.mean(dim=(-1, -2))     # [B,C] each


def _fused_ok(x, y, kernel_size, kernel_sigma):
    return (x.is_cuda and y.is_cuda and x.dtype == torch.float32 and kernel_size == 11 and kernel_sigma == 1.5
            and not y.requires_grad and x.size(0) * x.size(1) <= 65535)


def multi_scale_ssim(x, y, data_range=1.0, scale_weights=None, kernel_size=11, kernel_sigma=1.5, k1=0.01, k2=0.03):
    if scale_weights is None:
        scale_weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], device=x.device, dtype=x.dtype)
    else:
        scale_weights = torch.as_tensor(scale_weights, device=x.device, dtype=x.dtype)
        scale_weights = scale_weights / scale_weights.sum()
    levels = scale_weights.numel()
    min_size = (kernel_size - 1) * 2 ** (levels - 1) + 1
    if x.size(-1) < min_size or x.size(-2) < min_size:
        raise ValueError(f"Invalid size of the input images, expected at least {min_size}x{min_size}.")
    x = x / float(data_range)
    y = y / float(data_range)
    fused = _fused_ok(x, y, kernel_size, kernel_sigma)       # CUDA tensors: one fused kernel per scale (csrc/msssim.cu)
    win = None if fused else _gaussian_window(kernel_size, kernel_sigma, x.device, x.dtype).repeat(x.size(1), 1, 1, 1)
    c1, c2 = k1 ** 2, k2 ** 2
    terms = []
    ssim_last = None
    for level in range(levels):
        if level > 0:
            pad = max(x.shape[2] % 2, x.shape[3] % 2)
            x = F.avg_pool2d(F.pad(x, [pad, 0, pad, 0], mode="replicate"), kernel_size=2, padding=0)
            y = F.avg_pool2d(F.pad(y, [pad, 0, pad, 0], mode="replicate"), kernel_size=2, padding=0)
        if fused:
            from . import functional as F_sic
            ssim_last, cs = F_sic.ssim_stats(x, y, c1, c2)
        else:
            ssim_last, cs = _ssim_and_cs(x, y, win, c1, c2)
        terms.append(cs)
    stacked = torch.relu(torch.stack(terms[:-1] + [ssim_last], dim=0))          # [levels,B,C]
    powered = stacked ** scale_weights.view(-1, 1, 1)
    per_image = powered[0]
    for lvl in range(1, levels):              # explicit product: torch.prod's backward synchronises (zero counting),
        per_image = per_image * powered[lvl]  # which a CUDA-graph capture of the training step does not permit
    return per_image.mean(1).mean(0)
