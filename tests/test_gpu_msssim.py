"""N3: fused SSIM-statistics kernels (one forward + one backward launch per MS-SSIM scale) against the PyTorch statement
of the same algorithm in the oracle (oracle/torch_port.ssim_maps / multi_scale_ssim: depthwise conv2d chain, float64)."""
import numpy as np
import pytest
import torch

from oracle import torch_port as TP

pytestmark = pytest.mark.gpu


def _mods():
    from domain_specific_image_compression_b200 import functional as F
    from domain_specific_image_compression_b200 import losses
    return F, losses


@pytest.mark.parametrize("shape", [(2, 3, 64, 64), (1, 3, 41, 53), (3, 1, 11, 11), (2, 3, 128, 96)])
def test_ssim_stats_forward_backward_vs_conv_chain(shape):
    F, losses = _mods()
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    y = torch.rand(*shape, device="cuda", generator=g)
    x0 = (y + 0.15 * torch.randn(*shape, device="cuda", generator=g)).clamp(0, 1)
    gs, gc = torch.randn(shape[0], shape[1], device="cuda", generator=g), torch.randn(shape[0], shape[1], device="cuda", generator=g)
    xa = x0.clone().requires_grad_(True)
    ss, cs = F.ssim_stats(xa, y)
    ((ss * gs).sum() + (cs * gc).sum()).backward()
    xb = x0.double().clone().requires_grad_(True)
    win = TP.gaussian_window(shape[1], 11, 1.5, torch.float64, x0.device)
    ss_r, cs_r = (m.mean(dim=(-1, -2)) for m in TP.ssim_maps(xb, y.double(), win, 0.01 ** 2, 0.03 ** 2))
    ((ss_r * gs.double()).sum() + (cs_r * gc.double()).sum()).backward()
    assert float((ss.double() - ss_r).abs().max()) < 2e-6 and float((cs.double() - cs_r).abs().max()) < 2e-6
    assert float((xa.grad.double() - xb.grad).abs().max()) <= 2e-5 * float(xb.grad.abs().max()) + 1e-9
    with pytest.raises(ValueError):
        F.ssim_stats(torch.rand(1, 3, 10, 30, device="cuda"), torch.rand(1, 3, 10, 30, device="cuda"))


def test_multi_scale_ssim_kernel_vs_oracle_chain():
    F, losses = _mods()
    g = torch.Generator(device="cuda").manual_seed(0)
    y = torch.rand(4, 3, 256, 256, device="cuda", generator=g)
    x0 = (y + 0.1 * torch.randn(4, 3, 256, 256, device="cuda", generator=g))
    w = torch.tensor([0.3, 0.5, 0.2], device="cuda")
    xa = x0.clone().requires_grad_(True)
    n0 = F.launch_count
    va = losses.multi_scale_ssim(xa.clamp(0, 1), y, 1.0, w)
    va.backward()
    assert F.launch_count - n0 == 7                         # 3 scales x (1 forward + 1 backward kernel) + the combination kernel
    xb = x0.clone().requires_grad_(True)
    vb = TP.multi_scale_ssim(xb.clamp(0, 1), y, 1.0, w)      # the oracle's conv2d chain on the same device
    vb.backward()
    assert abs(float(va) - float(vb)) < 2e-6
    assert float((xa.grad - xb.grad).abs().max()) <= 1e-4 * float(xb.grad.abs().max()) + 1e-10
    # identical images -> 1, and the channels_last / odd-size (replicate-padded pooling) paths run
    assert abs(float(losses.multi_scale_ssim(y, y, 1.0, w)) - 1.0) < 1e-6
    z = torch.rand(2, 3, 101, 77, device="cuda", generator=g)
    a = losses.multi_scale_ssim(z.contiguous(memory_format=torch.channels_last), z * 0.9, 1.0, w)
    b = TP.multi_scale_ssim(z, z * 0.9, 1.0, w)
    assert abs(float(a) - float(b)) < 2e-6
    # float16 reconstructions (the reference's autocast training) are computed in float32 by the kernel
    c = losses.multi_scale_ssim(z.half(), z * 0.9, 1.0, w)
    assert abs(float(c) - float(b)) < 2e-3


@pytest.mark.parametrize("H,W", [(108, 84), (64, 64), (100, 76), (202, 44)])
def test_fused_pooling_between_scales_vs_oracle_chain(H, W):
    """Even sizes: the 2x2 average pooling between scales is written by the previous scale's forward kernel and its backward is added
    inside that scale's backward kernel (sic_ssim_fwd_pool / sic_ssim_bwd_pool).  Value and gradient vs the oracle's conv2d +
    avg_pool2d chain, at sizes that are not multiples of the 32-pixel tile (ragged last tiles own up to 42 rows / columns) and
    that turn odd at a coarser scale (100x76 -> 50x38 -> 25x19)."""
    F, losses = _mods()
    g = torch.Generator(device="cuda").manual_seed(H + W)
    y = torch.rand(2, 3, H, W, device="cuda", generator=g)
    x0 = (y + 0.1 * torch.randn(2, 3, H, W, device="cuda", generator=g))
    w = torch.tensor([0.3, 0.5, 0.2], device="cuda")
    xa = x0.clone().requires_grad_(True)
    va = losses.multi_scale_ssim(xa.clamp(0, 1), y, 1.0, w)
    va.backward()
    xb = x0.clone().requires_grad_(True)
    vb = TP.multi_scale_ssim(xb.clamp(0, 1), y, 1.0, w)
    vb.backward()
    assert abs(float(va) - float(vb)) < 2e-6
    assert float((xa.grad - xb.grad).abs().max()) <= 1e-4 * float(xb.grad.abs().max()) + 1e-10


@pytest.mark.parametrize("x_cl,y_cl", [(False, False), (True, True), (True, False), (False, True)])
@pytest.mark.parametrize("weights", [(0.3, 0.5, 0.2), None])
def test_fused_msssim_with_clamp_and_channels_last_vs_oracle_chain(x_cl, y_cl, weights):
    """The whole distortion term of model.py:96-101 in L + 1 + L launches: reconstruction with values outside [0, 1] (clamped inside the
    kernels, gradient zero there), either image planar or channels-last (what the fused first / last layers read and write), three
    scales with the reference's weights and piq's five default ones; value, gradient and gradient layout vs the oracle's chain."""
    F, losses = _mods()
    g = torch.Generator(device="cuda").manual_seed(11)
    B, H, W = 3, 256, 192
    y = torch.rand(B, 3, H, W, device="cuda", generator=g)
    x0 = y + 0.25 * torch.randn(B, 3, H, W, device="cuda", generator=g)          # ~15 % of the pixels leave [0, 1]
    w = None if weights is None else torch.tensor(weights, device="cuda")
    fmt = lambda t, cl: t.contiguous(memory_format=torch.channels_last) if cl else t.contiguous()
    xa = fmt(x0.clone(), x_cl).requires_grad_(True)
    n0 = F.launch_count
    va = losses.multi_scale_ssim(xa, fmt(y, y_cl), 1.0, w, clamp01=True)
    (va * 3.0).backward()
    L = 3 if weights else 5
    assert F.launch_count - n0 == 2 * L + 1
    xb = x0.double().clone().requires_grad_(True)
    vb = TP.multi_scale_ssim(xb.clamp(0, 1), y.double(), 1.0, None if w is None else w.double())
    (vb * 3.0).backward()
    assert abs(float(va) - float(vb)) < 3e-6
    assert xa.grad.is_contiguous(memory_format=torch.channels_last if x_cl else torch.contiguous_format)
    assert float((xa.grad.double() - xb.grad).abs().max()) <= 1e-4 * float(xb.grad.abs().max()) + 1e-10
    outside = (x0 < 0) | (x0 > 1)
    assert bool(outside.any()) and bool((xa.grad[outside] == 0).all())


def test_rate_distortion_loss_uses_the_fused_distortion():
    """rate_distortion_loss(dist='msssim') on a channels-last reconstruction: same loss and d loss / d x_hat as the reference's op order
    (clamp, MS-SSIM chain, 1 - ., lambda * D + R) in the oracle."""
    import domain_specific_image_compression_b200 as sic
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(2, 3, 128, 128, device="cuda", generator=g)
    xh0 = (x + 0.2 * torch.randn(2, 3, 128, 128, device="cuda", generator=g)).contiguous(memory_format=torch.channels_last)
    nll_y = torch.rand(2, 8, 8, 8, device="cuda", generator=g).requires_grad_(True)
    nll_z = torch.rand(2, 4, 2, 2, device="cuda", generator=g)
    xa = xh0.clone().requires_grad_(True)
    loss, R, D = sic.rate_distortion_loss({"x_hat": xa, "nll_y": nll_y, "nll_z": nll_z}, x, 100.0, "msssim")
    loss.backward()
    xb = xh0.double().clone().requires_grad_(True)
    w = torch.tensor([0.3, 0.5, 0.2], device="cuda", dtype=torch.float64)
    D_ref = 1.0 - TP.multi_scale_ssim(xb.clamp(0, 1), x.double(), 1.0, w)
    R_ref = (nll_y.detach().double().sum() + nll_z.double().sum()) / (2 * 128 * 128)
    (100.0 * D_ref + R_ref).backward()
    assert abs(float(D) - float(D_ref)) < 3e-6 and abs(float(R) - float(R_ref)) < 1e-6 * float(R_ref)
    assert abs(float(loss) - float(100.0 * D_ref + R_ref)) < 1e-5 * float(loss)
    assert not R.requires_grad and not D.requires_grad
    assert float((xa.grad.double() - xb.grad).abs().max()) <= 1e-4 * float(xb.grad.abs().max()) + 1e-10
    assert torch.allclose(nll_y.grad, torch.full_like(nll_y, 1.0 / (2 * 128 * 128)), rtol=1e-6, atol=0)
    # a negative total (cannot happen with real likelihoods; the reference clamps it, model.py:79): R = 0 and no gradient through it
    neg = (-nll_y.detach()).requires_grad_(True)
    loss2, R2, _ = sic.rate_distortion_loss({"x_hat": xh0, "nll_y": neg, "nll_z": nll_z * 0}, x, 100.0, "mse")
    loss2.backward()
    assert float(R2) == 0.0 and float(neg.grad.abs().max()) == 0.0
    assert abs(float(loss2) - 100.0 * float(torch.nn.functional.mse_loss(xh0, x))) <= 1e-6 * float(loss2)


def test_no_cpu_path_for_the_loss():
    """north_star: no CPU fallback.  The distortion term raises on CPU tensors like every other op of the package."""
    import domain_specific_image_compression_b200 as sic
    _, losses = _mods()
    w = torch.tensor([0.3, 0.5, 0.2])
    with pytest.raises(sic.SicError):
        losses.multi_scale_ssim(torch.rand(1, 3, 64, 64), torch.rand(1, 3, 64, 64), 1.0, w)
