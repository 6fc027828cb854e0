"""N2: the fused first analysis layer (conv 3->C 3x3 + bias + GDN in one tcgen05 kernel, csrc/conv0_gdn.cu) against the reference's
op chain (layers.py:49-51: `conv(3, N, 3, 1)` then `GDN(N)`, arithmetic layers.py:19-27) evaluated in float64 on the same device,
forward and backward.  Tolerance-only path: the convolution is evaluated to fp32 accuracy (exact tf32 hi/lo splits), not bit for
bit like cuDNN; eval/compress never use it."""
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

OFFSET = 2.0 ** -18


def _params(C, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.randn(C, 3, 3, 3, device="cuda", generator=g) * 0.3
    bias = torch.randn(C, device="cuda", generator=g) * 0.2
    beta = torch.sqrt(torch.rand(C, device="cuda", generator=g) + 0.5)
    gw = torch.sqrt(torch.rand(C, 1, 1, 1, device="cuda", generator=g) * 0.3 + 0.01)
    return w, bias, beta, gw


def _chain64(x, w, bias, beta, gw):
    """layers.py:49-51 in float64: conv -> + bias -> x / sqrt(beta_eff + gamma_eff x^2)."""
    v = TF.conv2d(x.double(), w.double(), None if bias is None else bias.double(), 1, 1)
    be = (beta.double() ** 2 - OFFSET).view(1, -1, 1, 1)
    ga = (gw.double().view(-1) ** 2 - OFFSET).view(1, -1, 1, 1)
    return v, v / torch.sqrt(be + ga * v * v)


SHAPES = [(1, 5, 7, 32), (2, 16, 16, 64), (1, 33, 20, 96), (2, 64, 64, 128), (1, 40, 24, 192), (3, 1, 1, 128), (1, 1, 300, 32),
          (2, 96, 160, 128), (1, 128, 100, 192)]


@pytest.mark.parametrize("B,H,W,C", SHAPES)
@pytest.mark.parametrize("with_bias", [True, False])
def test_forward_vs_float64_chain(B, H, W, C, with_bias):
    """Fewer positions than one tile, ragged tails, image borders on every side (1-pixel-high and 1-pixel-wide images), several
    tiles per persistent CTA (2 x 96 x 160 = 240 tiles on 148 SMs), both M-block layouts (C <= 128, C = 192)."""
    from domain_specific_image_compression_b200 import functional as F
    w, bias, beta, gw = _params(C, C + H)
    if not with_bias:
        bias = None
    x = torch.rand(B, 3, H, W, device="cuda")
    for xin in (x, x.contiguous(memory_format=torch.channels_last)):
        y, v = F.conv0_gdn(xin, w, bias, beta, gw, return_v=True)
        assert y.shape == (B, C, H, W) and y.is_contiguous(memory_format=torch.channels_last)
        v64, y64 = _chain64(x, w, bias, beta, gw)
        # exact operand splits + fp32 accumulation of 27 products: a few ulp of the largest partial sum
        assert float((v.double() - v64).abs().max()) <= 2e-6 * float(v64.abs().max()) + 1e-7
        assert float((y.double() - y64).abs().max()) <= 4e-6 * float(y64.abs().max()) + 1e-7
    y2 = F.conv0_gdn(x, w, bias, beta, gw)
    assert torch.equal(y2, y)                                    # deterministic, and the v output does not change y


@pytest.mark.parametrize("B,H,W,C", SHAPES)
def test_backward_vs_float64_autograd(B, H, W, C):
    """dW (the second tensor-core contraction, K = positions, dv read from tensor memory), d(bias), d(beta), d(gamma_conv.weight)
    against torch.autograd through the float64 chain."""
    from domain_specific_image_compression_b200 import functional as F
    w, bias, beta, gw = _params(C, 3 * C + W)
    x = torch.rand(B, 3, H, W, device="cuda")
    go = torch.randn(B, C, H, W, device="cuda")
    ps = [t.clone().requires_grad_(True) for t in (w, bias, beta, gw)]
    y = F.conv0_gdn(x, *ps)
    y.backward(go)
    p64 = [t.double().clone().requires_grad_(True) for t in (w, bias, beta, gw)]
    v = TF.conv2d(x.double(), p64[0], p64[1], 1, 1)
    be = (p64[2] ** 2 - OFFSET).view(1, -1, 1, 1)
    ga = (p64[3].view(-1) ** 2 - OFFSET).view(1, -1, 1, 1)
    (v / torch.sqrt(be + ga * v * v)).backward(go.double())
    for name, a, b in zip(("weight", "bias", "beta", "gamma_conv.weight"), ps, p64):
        assert a.grad is not None and a.grad.shape == a.shape, name
        err = float((a.grad.double() - b.grad).abs().max())
        # fp32 per-lane partial sums over up to 31 k positions, then binary64 folds: ~1e-6 relative to the largest gradient entry
        assert err <= 2e-5 * float(b.grad.abs().max()) + 1e-6, (name, err, float(b.grad.abs().max()))
    # deterministic: a second backward gives the same bits
    ps2 = [t.clone().requires_grad_(True) for t in (w, bias, beta, gw)]
    F.conv0_gdn(x, *ps2).backward(go)
    for a, b in zip(ps, ps2):
        assert torch.equal(a.grad, b.grad)


def test_full_site_properties():
    """BASELINE.json cfg2's site (16 x 128 x 256 x 256, the size the roofline is quoted on) through size-independent properties:
    the per-channel sums of y^2-weighted identities cannot be checked cheaply in float64 at this size, so (i) a 1/16 slice of the
    batch must reproduce the matching slice of the full result bit for bit (tiles never straddle semantics: per-position results
    do not depend on the tile or CTA that computed them), and (ii) the gradients of the full batch equal the fixed-order sum over
    slices to fp32 accumulation accuracy."""
    from domain_specific_image_compression_b200 import functional as F
    B, H, W, C = 16, 256, 256, 128
    w, bias, beta, gw = _params(C, 7)
    x = torch.rand(B, 3, H, W, device="cuda")
    go = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    ps = [t.clone().requires_grad_(True) for t in (w, bias, beta, gw)]
    y = F.conv0_gdn(x, *ps)
    y.backward(go)
    acc = [torch.zeros_like(t, dtype=torch.float64) for t in ps]
    for b in (0, 7, 15):
        yb = F.conv0_gdn(x[b:b + 1], w, bias, beta, gw)
        assert torch.equal(yb, y[b:b + 1])
    for b in range(B):
        q = [t.clone().requires_grad_(True) for t in (w, bias, beta, gw)]
        F.conv0_gdn(x[b:b + 1], *q).backward(go[b:b + 1])
        for a, t in zip(acc, q):
            a += t.grad.double()
    for name, a, t in zip(("weight", "bias", "beta", "gamma_conv.weight"), acc, ps):
        assert float((t.grad.double() - a).abs().max()) <= 2e-5 * float(a.abs().max()) + 1e-6, name


def test_errors():
    from domain_specific_image_compression_b200 import functional as F
    from domain_specific_image_compression_b200._lib import SicError
    w, bias, beta, gw = _params(64, 1)
    x = torch.rand(1, 3, 8, 8, device="cuda")
    with pytest.raises(SicError):
        F.conv0_gdn(x.cpu(), w, bias, beta, gw)                                     # no CPU path
    with pytest.raises(SicError):
        F.conv0_gdn(x, torch.randn(48, 3, 3, 3, device="cuda"), None, torch.ones(48, device="cuda"), torch.ones(48, device="cuda"))
    with pytest.raises(SicError):
        F.conv0_gdn(x.clone().requires_grad_(True), w, bias, beta, gw)              # the image gets no gradient
    with pytest.raises(SicError):
        F.conv0_gdn(x, w, bias, beta[:32], gw)


def test_model_training_step_with_fused_first_layer():
    """The whole model in training mode with layers.FUSE_FIRST_LAYER on vs off: same loss and gradients to conv-rounding accuracy
    (the unfused first layer is a cuDNN fp32 convolution); eval mode ignores the switch, so the bit-exact latents of the reference
    path are untouched."""
    import domain_specific_image_compression_b200 as sic
    from domain_specific_image_compression_b200 import layers as L
    torch.manual_seed(5)
    m = sic.CompressionModel(N=32, M=48, spatial_params=False, min_nu=2.0, max_nu=100.0).cuda()
    with torch.no_grad():
        m.g_a.g_a[14].weight.mul_(40.0)
        m.h_a.h_a[6].weight.mul_(40.0)
    x = torch.rand(2, 3, 64, 80, device="cuda")
    ny, nz = torch.rand(2, 48, 4, 5, device="cuda") - 0.5, torch.rand(2, 32, 1, 2, device="cuda") - 0.5
    res = {}
    old_tf32, old_ba = torch.backends.cudnn.allow_tf32, L.FUSE_BIAS_ACT
    torch.backends.cudnn.allow_tf32 = False
    L.FUSE_BIAS_ACT = False          # the fused layer hands on channels-last activations, where the bias kernels would add launches of their own
    try:
        for fused in (False, True):
            L.FUSE_FIRST_LAYER = fused
            m.train()
            m.zero_grad(set_to_none=True)
            n0 = sic.functional.launch_count
            out = m(x, "noise", noise_y=ny, noise_z=nz)
            loss, _, _ = sic.rate_distortion_loss(out, x, 100.0, "mse")
            loss.backward()
            res[fused] = (float(loss), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}, out["y"].clone(),
                          sic.functional.launch_count - n0)
        m.eval()
        with torch.no_grad():
            e_on = m(x, "round")["y_tilde"]
            L.FUSE_FIRST_LAYER = False
            e_off = m(x, "round")["y_tilde"]
        assert torch.equal(e_on, e_off)
    finally:
        L.FUSE_FIRST_LAYER = False
        L.FUSE_BIAS_ACT = old_ba
        torch.backends.cudnn.allow_tf32 = old_tf32
    (l0, g0, y0, n0), (l1, g1, y1, n1) = res[False], res[True]
    assert n1 == n0                                               # our launch count is unchanged (1 fwd + 2 bwd either way); two cuDNN launches disappear
    assert abs(l0 - l1) <= 1e-4 * abs(l0)
    assert float((y0 - y1).abs().max()) <= 1e-3 * float(y0.abs().max())
    assert g0.keys() == g1.keys()
    for n in g0:
        assert float((g0[n] - g1[n]).abs().max()) <= 2e-3 * float(g0[n].abs().max()) + 1e-7, n
