"""Plain-PyTorch restatement of the reference model's forward / loss / training step (TEST INFRASTRUCTURE).

Purpose: (1) the CPU baseline of bench.py (`cpu_baseline`, `--impl reference`): the reference's own eager op chains on the
box's host cores — /root/reference is not present on the GPU box, so the port travels instead; (2) on the GPU box, the
"reference on B200 in eager PyTorch" comparator for bit-exactness of GDN and the latents.  It is written functionally
over a state_dict with the reference's key names, op for op in the reference's order:

    GDN            layers.py:19-27          analysis/synthesis/hyper stacks  layers.py:49-73, 81-98, 107-113, 122-152
    quantize       model.py:27-35           forward                          model.py:37-72
    Student-t nll  distributions.py:20-31   Gaussian nll                     distributions.py:39-46
    loss           model.py:75-107          MS-SSIM (piq, restated)          model.py:96-101

Pinned against the imported reference by tests/golden/model_small.npz (same weights, same input, same noise).
Never imported by the product package.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

LOG2E = 1.0 / math.log(2.0)
OFFSET = 2 ** -18


def gdn(x, beta_param, weight, inverse):
    beta = beta_param ** 2 - OFFSET
    gamma = weight ** 2 - OFFSET
    denom = torch.sqrt(beta.view(1, -1, 1, 1) + F.conv2d(x ** 2, gamma, bias=None, groups=x.size(1)))
    return x * denom if inverse else x / denom


def _conv(sd, key, x, stride, k):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=stride, padding=(k - 1) // 2)


def _deconv(sd, key, x):
    return F.conv_transpose2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=2, padding=2, output_padding=1)


def _gdn_site(sd, key, x, inverse):
    return gdn(x, sd[key + ".beta"], sd[key + ".gamma_conv.weight"], inverse)


def analysis(sd, x):
    p = "g_a.g_a."
    t = _gdn_site(sd, p + "1", _conv(sd, p + "0", x, 1, 3), False)
    i = 2
    for _ in range(3):
        t = _gdn_site(sd, p + str(i + 1), _conv(sd, p + str(i), t, 2, 5), False)
        t = _gdn_site(sd, p + str(i + 3), _conv(sd, p + str(i + 2), t, 1, 3), False)
        i += 4
    return _conv(sd, p + "14", t, 2, 5)


def synthesis(sd, y_hat):
    p = "g_s.g_s."
    t = y_hat
    i = 0
    for _ in range(3):
        t = _gdn_site(sd, p + str(i + 1), _deconv(sd, p + str(i), t), True)
        t = _gdn_site(sd, p + str(i + 3), _conv(sd, p + str(i + 2), t, 1, 3), True)
        i += 4
    return _deconv(sd, p + "12", t)


def hyper_analysis(sd, y):
    p = "h_a.h_a."
    t = F.relu(_conv(sd, p + "0", y, 1, 3))
    t = F.relu(_conv(sd, p + "2", t, 1, 3))
    t = F.relu(_conv(sd, p + "4", t, 2, 5))
    return _conv(sd, p + "6", t, 2, 5)


def hyper_synthesis(sd, z, spatial_params=False):
    t = F.relu(_deconv(sd, "h_s.h_s.0", z))
    t = F.relu(_deconv(sd, "h_s.h_s.2", t))
    if spatial_params:
        return _conv(sd, "h_s.to_sigma", t, 1, 3), _conv(sd, "h_s.to_nu", t, 1, 3)
    return hyper_heads(sd, t)


def hyper_heads(sd, t):
    """layers.py:146-151: pool -> mlp_sigma / mlp_nu -> expand over (h,w)."""
    pooled = F.adaptive_avg_pool2d(t, 1)
    heads = []
    for name in ("h_s.mlp_sigma", "h_s.mlp_nu"):
        u = F.relu(F.conv2d(pooled, sd[name + ".0.weight"], sd[name + ".0.bias"]))
        u = F.conv2d(u, sd[name + ".2.weight"], sd[name + ".2.bias"])
        heads.append(u.expand(-1, -1, t.size(2), t.size(3)))
    return heads[0], heads[1]


def hyper_tail(sd, t, min_nu=2.0, max_nu=100.0):
    """What follows the two transposed convolutions of h_s when spatial_params=False: layers.py:146-151 then model.py:54-55.
    Returns (sigma, nu) as [B,M,1,1]."""
    log_sigma, log_nu = hyper_heads(sd, t)
    sigma = torch.exp(log_sigma).mean(dim=(2, 3), keepdim=True)
    nu = torch.clamp(torch.exp(log_nu).mean(dim=(2, 3), keepdim=True), min_nu, max_nu)
    return sigma, nu


def studentt_nll(x, sigma, nu):
    sigma = torch.clamp(sigma, min=1e-3, max=1e3)
    nu = torch.clamp(nu, min=2.0, max=100.0)
    logC = torch.lgamma((nu + 1.0) / 2.0) - torch.lgamma(nu / 2.0) - 0.5 * torch.log(nu * torch.pi) - torch.log(sigma)
    quad = (x / sigma) ** 2
    return -(logC - ((nu + 1.0) / 2.0) * torch.log1p(quad / nu)) * LOG2E


def gaussian_nll(x, log_sigma):
    sigma = torch.clamp(torch.exp(log_sigma).view(1, -1, 1, 1), min=1e-3, max=1e3)
    var = sigma ** 2
    return -(-0.5 * torch.log(2 * torch.pi * var) - 0.5 * (x ** 2) / var) * LOG2E


def quantize(x, mode, noise=None):
    if mode == "noise":
        return x + (noise if noise is not None else torch.empty_like(x).uniform_(-0.5, 0.5))
    if mode == "round":
        return torch.round(x)
    raise ValueError(f"Unknown quant mode: {mode}")


def forward(sd: Dict[str, torch.Tensor], x, quant_mode="noise", training=True, spatial_params=False, min_nu=2.0, max_nu=100.0,
            noise_y=None, noise_z=None):
    y = analysis(sd, x)
    z = hyper_analysis(sd, y)
    y_tilde = quantize(y, quant_mode, noise_y)
    z_tilde = quantize(z, quant_mode, noise_z)
    log_sigma, log_nu = hyper_synthesis(sd, z_tilde, spatial_params)
    if spatial_params:
        sigma = torch.exp(log_sigma)
        nu = torch.clamp(torch.exp(log_nu), min=min_nu, max=max_nu)
    else:
        sigma = torch.exp(log_sigma).mean(dim=(2, 3), keepdim=True).expand_as(y_tilde)
        nu = torch.clamp(torch.exp(log_nu).mean(dim=(2, 3), keepdim=True), min_nu, max_nu).expand_as(y_tilde)
    nll_y = studentt_nll(y_tilde, sigma, nu)
    nll_z = gaussian_nll(z_tilde, sd["z_prior.log_sigma"])
    y_hat = y_tilde if training else torch.round(y)
    return {"x_hat": synthesis(sd, y_hat), "nll_y": nll_y, "nll_z": nll_z, "y": y, "y_tilde": y_tilde, "z": z,
            "z_tilde": z_tilde, "sigma": sigma, "nu": nu}


def ssim_maps(x, y, win, c1, c2):
    """SSIM map and its contrast-structure factor for one scale: the depthwise "valid" conv2d chain of piq's _ssim_per_channel."""
    C = x.size(1)
    mu_x, mu_y = F.conv2d(x, win, groups=C), F.conv2d(y, win, groups=C)
    mu_xx, mu_yy, mu_xy = mu_x * mu_x, mu_y * mu_y, mu_x * mu_y
    s_xx = F.conv2d(x * x, win, groups=C) - mu_xx
    s_yy = F.conv2d(y * y, win, groups=C) - mu_yy
    s_xy = F.conv2d(x * y, win, groups=C) - mu_xy
    cs_map = (2.0 * s_xy + c2) / (s_xx + s_yy + c2)
    ss_map = (2.0 * mu_xy + c1) / (mu_xx + mu_yy + c1) * cs_map
    return ss_map, cs_map


def gaussian_window(channels, size=11, sigma=1.5, dtype=torch.float32, device="cpu"):
    coords = torch.arange(size, dtype=dtype, device=device) - (size - 1) / 2.0
    g1 = torch.exp(-(coords ** 2) / (2.0 * sigma ** 2))
    g1 = g1 / g1.sum()
    return (g1[:, None] * g1[None, :])[None, None].repeat(channels, 1, 1, 1)


def multi_scale_ssim(x, y, data_range=1.0, scale_weights=None, kernel_size=11, kernel_sigma=1.5, k1=0.01, k2=0.03):
    """piq.multi_scale_ssim as the reference calls it (model.py:96-101; piq 0.8.0, Requirements.txt).  piq is neither vendored
    with the reference nor installed here, so this restates its published algorithm (PARITY UNPINNED): 11x11 Gaussian window
    (sigma 1.5) as a depthwise "valid" convolution, c1 = (k1 L)^2, c2 = (k2 L)^2, 2x2 average pooling between scales
    (replicate-padding odd sizes on the top/left), relu on the per-scale terms, prod_i cs_i^w_i * ssim_last^w_last with the
    weights normalised to sum 1, mean over channels then over the batch.  Plain eager PyTorch: this is the oracle's own copy, so
    that the CPU baseline arm runs none of the product's code."""
    if scale_weights is None:
        scale_weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], device=x.device, dtype=x.dtype)
    else:
        scale_weights = torch.as_tensor(scale_weights, device=x.device, dtype=x.dtype)
        scale_weights = scale_weights / scale_weights.sum()
    levels = scale_weights.numel()
    min_size = (kernel_size - 1) * 2 ** (levels - 1) + 1
    if x.size(-1) < min_size or x.size(-2) < min_size:
        raise ValueError(f"Invalid size of the input images, expected at least {min_size}x{min_size}.")
    x, y = x / float(data_range), y / float(data_range)
    coords = torch.arange(kernel_size, dtype=x.dtype, device=x.device) - (kernel_size - 1) / 2.0
    g1 = torch.exp(-(coords ** 2) / (2.0 * kernel_sigma ** 2))
    g1 = g1 / g1.sum()
    C = x.size(1)
    win = (g1[:, None] * g1[None, :])[None, None].repeat(C, 1, 1, 1)
    c1, c2 = k1 ** 2, k2 ** 2
    terms = []
    for level in range(levels):
        if level > 0:
            pad = max(x.shape[2] % 2, x.shape[3] % 2)
            x = F.avg_pool2d(F.pad(x, [pad, 0, pad, 0], mode="replicate"), kernel_size=2, padding=0)
            y = F.avg_pool2d(F.pad(y, [pad, 0, pad, 0], mode="replicate"), kernel_size=2, padding=0)
        ss_map, cs_map = ssim_maps(x, y, win, c1, c2)
        terms.append((ss_map if level == levels - 1 else cs_map).mean(dim=(-1, -2)))     # [B,C]
    stacked = torch.relu(torch.stack(terms, dim=0)) ** scale_weights.view(-1, 1, 1)
    return torch.prod(stacked, dim=0).mean(1).mean(0)


def loss_fn(out, x, lambda_rd=10000.0, dist="msssim", msssim=None):
    msssim = msssim or multi_scale_ssim
    n, _, h, w = x.shape
    R = torch.clamp((out["nll_y"].sum() + out["nll_z"].sum()) / (n * h * w), min=0.0)
    if dist == "mse":
        D = F.mse_loss(out["x_hat"], x)
    elif dist == "msssim":
        D = 1.0 - msssim(out["x_hat"].clamp(0, 1), x, data_range=1.0, scale_weights=torch.tensor([0.3, 0.5, 0.2], device=x.device))
    else:
        raise ValueError("dist must be 'mse' or 'msssim'")
    return lambda_rd * D + R, R.detach(), D.detach()


def init_state(N=128, M=192, seed=42, device="cpu", spread=True) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's parameter shapes and (torch-default-like) scales; 'spread' widens the last
    analysis convs so the latents are not degenerate."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, cout, cin, k, transposed=False):
        fan_in = (cout if transposed else cin) * k * k
        bound = 1.0 / math.sqrt(fan_in)
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        sd[key + ".weight"] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[key + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def gdn_site(key, c):
        sd[key + ".beta"] = torch.sqrt(torch.ones(c) + OFFSET)
        sd[key + ".gamma"] = torch.sqrt(torch.eye(c) * 0.1 + OFFSET)
        sd[key + ".gamma_conv.weight"] = sd[key + ".gamma"].diag().view(c, 1, 1, 1).clone()

    conv("g_a.g_a.0", N, 3, 3)
    gdn_site("g_a.g_a.1", N)
    i = 2
    for _ in range(3):
        conv(f"g_a.g_a.{i}", N, N, 5); gdn_site(f"g_a.g_a.{i+1}", N)
        conv(f"g_a.g_a.{i+2}", N, N, 3); gdn_site(f"g_a.g_a.{i+3}", N)
        i += 4
    conv("g_a.g_a.14", M, N, 5)
    conv("g_s.g_s.0", N, M, 5, True); gdn_site("g_s.g_s.1", N)
    conv("g_s.g_s.2", N, N, 3); gdn_site("g_s.g_s.3", N)
    for i in (4, 8):
        conv(f"g_s.g_s.{i}", N, N, 5, True); gdn_site(f"g_s.g_s.{i+1}", N)
        conv(f"g_s.g_s.{i+2}", N, N, 3); gdn_site(f"g_s.g_s.{i+3}", N)
    conv("g_s.g_s.12", 3, N, 5, True)
    conv("h_a.h_a.0", N, M, 3); conv("h_a.h_a.2", N, N, 3); conv("h_a.h_a.4", N, N, 5); conv("h_a.h_a.6", N, N, 5)
    conv("h_s.h_s.0", N, N, 5, True); conv("h_s.h_s.2", N, N, 5, True)
    for name in ("h_s.mlp_sigma", "h_s.mlp_nu"):
        conv(name + ".0", N, N, 1); conv(name + ".2", M, N, 1)
    sd["z_prior.log_sigma"] = torch.zeros(N)
    if spread:
        sd["g_a.g_a.14.weight"] *= 40.0
        sd["h_a.h_a.6.weight"] *= 40.0
        sd["h_s.mlp_nu.2.bias"] += 1.5
    return {k: v.to(device) for k, v in sd.items()}


def train_step(sd, opt, x, lambda_rd=10000.0, dist="mse", msssim=None, grad_clip=1.0):
    """train.py:193-204 without AMP (BASELINE configs are fp32): zero_grad, forward 'noise', loss, backward, clip, Adam."""
    opt.zero_grad(set_to_none=True)
    out = forward(sd, x, "noise", training=True)
    loss, R, D = loss_fn(out, x, lambda_rd, dist, msssim)
    loss.backward()
    params = [p for p in sd.values() if p.requires_grad]
    if grad_clip > 0:
        torch.nn.utils.clip_grad_norm_(params, grad_clip)
    opt.step()
    return loss.detach()
