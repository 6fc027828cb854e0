"""Bias add (+ ReLU) after a bias-free convolution and its adjoint on channels-last activations (csrc/bias_act.cu) against PyTorch's own
op chain - conv -> add_(bias) -> relu_ forward (bit-identical values), threshold_backward + sum over (B, H, W) backward - which is what
the reference's hyper transforms (layers.py:104-139) and last analysis convolution (layers.py:73) run."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,C,H,W", [(2, 128, 16, 16), (1, 192, 5, 7), (3, 320, 8, 8), (2, 3, 64, 48), (1, 128, 1, 1), (16, 128, 16, 16), (1, 5, 3, 3)])
@pytest.mark.parametrize("relu", [False, True])
def test_bias_act_forward_bit_exact_and_backward(B, C, H, W, relu):
    from domain_specific_image_compression_b200 import functional as F
    gen = torch.Generator(device="cuda").manual_seed(B * 1000 + C)
    t0 = torch.randn(B, C, H, W, device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
    b0 = torch.randn(C, device="cuda", generator=gen)
    t0[:, :, 0, 0] = -b0                                            # pre-activations of exactly zero: the ReLU mask is `> 0`
    g = torch.randn(B, C, H, W, device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
    # ours: the op works in place on a non-leaf (the convolution's output), so feed it one
    src = t0.clone().requires_grad_(True)
    b = b0.clone().requires_grad_(True)
    y = F.bias_act(src * 1.0, b, relu=relu)
    y.backward(g)
    # PyTorch's chain
    src_r, b_r = t0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    y_r = src_r * 1.0 + b_r.view(1, -1, 1, 1)
    if relu:
        y_r = torch.relu_(y_r)
    y_r.backward(g)
    assert torch.equal(y.detach().view(torch.int32), y_r.detach().view(torch.int32))      # bit pattern: signed zeros included
    assert torch.equal(src.grad, src_r.grad)
    ref = src_r.grad.double().sum(dim=(0, 2, 3))
    assert float((b.grad.double() - ref).abs().max()) <= 2e-6 * float(src_r.grad.double().abs().sum(dim=(0, 2, 3)).max()) + 1e-12
    assert float((b.grad - b_r.grad).abs().max()) <= 1e-5 * float(b_r.grad.abs().max()) + 1e-6


def test_channel_sum_and_refusals():
    import domain_specific_image_compression_b200 as sic
    from domain_specific_image_compression_b200 import functional as F
    gen = torch.Generator(device="cuda").manual_seed(4)
    for shape in ((16, 3, 256, 256), (2, 192, 16, 16), (1, 1024, 2, 3)):
        g = torch.randn(*shape, device="cuda", generator=gen)
        for t in (g, g.contiguous(memory_format=torch.channels_last)):
            s = F.channel_sum(t)
            ref = g.double().sum(dim=(0, 2, 3))
            assert float((s.double() - ref).abs().max()) <= 2e-6 * float(g.double().abs().sum(dim=(0, 2, 3)).max())
            assert torch.equal(s, F.channel_sum(t))                                       # deterministic
    x = torch.randn(2, 8, 4, 4, device="cuda")
    assert not F.bias_act_supported(x, torch.zeros(8, device="cuda"))                     # NCHW: the caller keeps PyTorch's add
    with pytest.raises(sic.SicError):
        F.bias_act(x, torch.zeros(8, device="cuda"))
    with pytest.raises(sic.SicError):
        F.bias_act(x.cpu(), torch.zeros(8))


def test_model_step_with_and_without_fused_bias_act():
    """Channels-last model, training mode: layers.FUSE_BIAS_ACT on vs off - identical forward values (same fp32 add, same ReLU), so
    the same loss bit for bit; gradients equal up to the summation order of d(bias)."""
    import domain_specific_image_compression_b200 as sic
    from domain_specific_image_compression_b200 import layers as L
    torch.manual_seed(3)
    m = sic.CompressionModel(N=32, M=48, spatial_params=False, min_nu=2.0, max_nu=100.0).cuda().to(memory_format=torch.channels_last)
    x = torch.rand(2, 3, 64, 64, device="cuda").contiguous(memory_format=torch.channels_last)
    ny, nz = torch.rand(2, 48, 4, 4, device="cuda") - 0.5, torch.rand(2, 32, 1, 1, device="cuda") - 0.5
    res = {}
    old = L.FUSE_BIAS_ACT
    try:
        with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
            for on in (False, True):
                L.FUSE_BIAS_ACT = on
                m.train()
                m.zero_grad(set_to_none=True)
                out = m(x, "noise", noise_y=ny, noise_z=nz)
                loss, _, _ = sic.rate_distortion_loss(out, x, 100.0, "msssim")
                loss.backward()
                res[on] = (loss.detach().clone(), out["z"].detach().clone(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
    finally:
        L.FUSE_BIAS_ACT = old
    (l0, z0, g0), (l1, z1, g1) = res[False], res[True]
    assert torch.equal(z0, z1) and torch.equal(l0, l1)
    assert g0.keys() == g1.keys()
    for n in g0:
        assert float((g0[n] - g1[n]).abs().max()) <= 1e-5 * float(g0[n].abs().max()) + 1e-9, n
