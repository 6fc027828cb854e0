"""Distortion terms of rate_distortion_loss (R2 in SURVEY.md 8(a)).  The per-scale SSIM statistics run in the fused CUDA kernels
(csrc/msssim.cu, row N3); pooling, relu, powers and the product over scales stay in torch (a few [B,3] tensors).  Like every
other op of the package there is NO CPU or eager path: CPU tensors raise SicError (the conv2d-chain statement of the same
algorithm lives in oracle/torch_port.py, test infrastructure only).

`multi_scale_ssim` restates piq 0.8.0's `multi_scale_ssim` (called at /root/reference/code/modelv2/model.py:96-101 with
data_range=1, scale_weights=[.3,.5,.2]); piq is a third-party package that is neither vendored with the reference nor
installed here, so this follows its published algorithm: Gaussian 11x11 window (sigma 1.5), k1=.01, k2=.03, "valid"
depthwise convolutions, 2x2 average pooling between scales (replicate-padding odd sizes on the top/left), relu on the
per-scale terms, prod_i cs_i^w_i * ssim_last^w_last, mean over channels then batch.  PARITY UNPINNED.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from . import functional as F_sic


def multi_scale_ssim(x, y, data_range=1.0, scale_weights=None, kernel_size=11, kernel_sigma=1.5, k1=0.01, k2=0.03, clamp01=False):
    """clamp01=True folds the x.clamp(0, 1) the reference applies to the reconstruction first (model.py:98) into the kernels."""
    normalize = scale_weights is not None      # piq normalises user-supplied weights, not its defaults
    if scale_weights is None:
        raw_weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], device=x.device, dtype=torch.float32)
    else:
        raw_weights = torch.as_tensor(scale_weights, device=x.device, dtype=torch.float32)
    levels = raw_weights.numel()
    min_size = (kernel_size - 1) * 2 ** (levels - 1) + 1
    if x.size(-1) < min_size or x.size(-2) < min_size:
        raise ValueError(f"Invalid size of the input images, expected at least {min_size}x{min_size}.")
    if kernel_size != 11 or kernel_sigma != 1.5:
        raise _lib.SicError("multi_scale_ssim: the kernel fixes the 11x11 / sigma 1.5 window the reference uses (model.py:96-101)")
    if y.requires_grad:
        raise _lib.SicError("multi_scale_ssim: differentiable w.r.t. the reconstruction only (the target must not require grad)")
    # float16 reconstructions arrive here under the reference's autocast training (train.py:196-199): computed in float32
    x, y = x.float(), y.float()
    if float(data_range) != 1.0:              # the reference calls with data_range = 1: no scaling pass (forward and backward) at all
        if clamp01:                           # the clamp applies to the unscaled reconstruction
            x, clamp01 = x.clamp(0, 1), False
        x = x / float(data_range)
        y = y / float(data_range)
    c1, c2 = k1 ** 2, k2 ** 2
    if levels <= 8 and all(((x.size(-2) >> l) | (x.size(-1) >> l)) & 1 == 0 for l in range(levels - 1)):
        # every pooled scale has even sizes (the training patches): all scales, their combination and the clamp in L + 1 launches
        return F_sic.msssim_fused(x, y, raw_weights, normalize, c1, c2, clamp01=clamp01)
    if clamp01:
        x = x.clamp(0, 1)
    scale_weights = raw_weights / raw_weights.sum() if normalize else raw_weights
    terms = []
    ssim_last = None
    pooled = None
    for level in range(levels):
        if level > 0:
            if pooled is not None:            # written by the previous scale's kernel (even sizes)
                x, y = pooled
            else:
                pad = max(x.shape[2] % 2, x.shape[3] % 2)
                if pad:                       # odd sizes: replicate-pad top/left, then pool, in torch
                    x = F.pad(x, [pad, 0, pad, 0], mode="replicate")
                    y = F.pad(y, [pad, 0, pad, 0], mode="replicate")
                x = F.avg_pool2d(x, kernel_size=2, padding=0)
                y = F.avg_pool2d(y, kernel_size=2, padding=0)
        fuse_pool = level + 1 < levels and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0
        if fuse_pool:                         # one kernel: this scale's statistics + the next scale's inputs
            ssim_last, cs, xp, yp = F_sic.ssim_stats(x, y, c1, c2, pool=True)
            pooled = (xp, yp)
        else:
            ssim_last, cs = F_sic.ssim_stats(x, y, c1, c2)       # raises on CPU tensors
            pooled = None
        terms.append(cs)
    stacked = torch.relu(torch.stack(terms[:-1] + [ssim_last], dim=0))          # [levels,B,C]
    powered = stacked ** scale_weights.view(-1, 1, 1)
    per_image = powered[0]
    for lvl in range(1, levels):              # explicit product: torch.prod's backward synchronises (zero counting),
        per_image = per_image * powered[lvl]  # which a CUDA-graph capture of the training step does not permit
    return per_image.mean(1).mean(0)
