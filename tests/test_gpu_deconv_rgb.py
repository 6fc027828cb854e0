"""N2, synthesis side: the last layer deconv(N, 3) (layers.py:96-98: ConvTranspose2d(N, 3, 5, stride 2, padding 2, output_padding 1)) as
library GEMM + the col2im / im2col gather kernels (csrc/deconv_rgb.cu), against torch's conv_transpose2d in float64, forward and
backward.  fp32 GEMMs here (TF32 off, as in every parity test); the gathers themselves are exact (pure data movement + fixed-order
fp32 sums of at most 9 taps)."""
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


SHAPES = [(1, 16, 1, 1), (2, 32, 3, 5), (1, 128, 16, 16), (3, 64, 7, 33), (2, 192, 20, 12), (1, 128, 64, 64)]


@pytest.mark.parametrize("B,N,H,W", SHAPES)
@pytest.mark.parametrize("with_bias", [True, False])
def test_forward_backward_vs_float64(B, N, H, W, with_bias):
    from domain_specific_image_compression_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(N + H)
    a0 = torch.randn(B, N, H, W, device="cuda", generator=g)
    w0 = torch.randn(N, 3, 5, 5, device="cuda", generator=g) * 0.1
    b0 = torch.randn(3, device="cuda", generator=g) if with_bias else None
    go = torch.randn(B, 3, 2 * H, 2 * W, device="cuda", generator=g)
    for fmt in (torch.contiguous_format, torch.channels_last):
        a = a0.clone(memory_format=fmt).requires_grad_(True)
        w = w0.clone().requires_grad_(True)
        b = None if b0 is None else b0.clone().requires_grad_(True)
        y = F.deconv_rgb(a, w, b)
        assert y.shape == (B, 3, 2 * H, 2 * W)
        y.backward(go)
        a64, w64 = a0.double().requires_grad_(True), w0.double().requires_grad_(True)
        b64 = None if b0 is None else b0.double().requires_grad_(True)
        y64 = TF.conv_transpose2d(a64, w64, b64, 2, 2, 1)
        y64.backward(go.double())
        tol = lambda ref: 2e-5 * float(ref.abs().max()) + 1e-6
        assert float((y.double() - y64).abs().max()) <= tol(y64)
        assert a.grad.shape == a.shape and float((a.grad.double() - a64.grad).abs().max()) <= tol(a64.grad)
        assert float((w.grad.double() - w64.grad).abs().max()) <= tol(w64.grad)
        if b0 is not None:
            assert float((b.grad.double() - b64.grad).abs().max()) <= tol(b64.grad)


def test_gathers_are_exact_adjoints():
    """col2im and im2col alone (no GEMM): col2im of a random D equals the float64 scatter definition to fp32 summation accuracy, and
    <col2im(D), g> == <D, im2col(g)> (they are adjoint linear maps: every tap element of D is used exactly once).  Shapes cover
    several tiles in both directions with ragged edges (tile = 8 x 32 positions)."""
    import ctypes
    from domain_specific_image_compression_b200 import _lib
    lib = _lib.load()
    for B, H, W in ((2, 9, 13), (1, 17, 70), (3, 1, 1)):
        P = B * H * W
        gen = torch.Generator(device="cuda").manual_seed(3 + H)
        D = torch.randn(P, 80, device="cuda", generator=gen)
        g = torch.randn(B, 2 * H, 2 * W, 3, device="cuda", generator=gen)
        out = torch.full((B, 2 * H, 2 * W, 3), float("nan"), device="cuda")
        dD = torch.full((P, 80), float("nan"), device="cuda")
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        assert lib.sic_deconv_rgb_col2im(vp(D), None, B, H, W, vp(out), st) == 0, lib.sic_last_error()
        assert lib.sic_deconv_rgb_im2col(vp(g), B, H, W, vp(dD), st) == 0, lib.sic_last_error()
        # definition: scatter D[p, m] to out[b, 2iy-2+kh, 2ix-2+kw, co]
        ref = torch.zeros(B, 2 * H + 4, 2 * W + 4, 3, dtype=torch.float64, device="cuda")      # padded by 2 on each side
        Dv = D[:, :75].double().view(B, H, W, 5, 5, 3)
        for kh in range(5):
            for kw in range(5):
                ref[:, kh:kh + 2 * H:2, kw:kw + 2 * W:2, :] += Dv[:, :, :, kh, kw, :]
        ref = ref[:, 2:2 + 2 * H, 2:2 + 2 * W, :]
        assert float((out.double() - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
        assert bool((dD[:, 75:] == 0).all())
        lhs = float((out.double() * g.double()).sum())
        rhs = float((D[:, :75].double() * dD[:, :75].double()).sum())
        assert abs(lhs - rhs) <= 1e-6 * abs(lhs) + 1e-6
    assert lib.sic_deconv_rgb_col2im(vp(D), None, 0, H, W, vp(out), st) == -1            # SIC_E_BADARG


def test_model_training_step_with_gemm_last_layer():
    """Whole model, training mode, layers.FAST_LAST_LAYER on vs off (cuDNN fp32): same loss, x_hat and gradients to fp32 conv
    accuracy; eval ignores the switch (bit-exact x_hat of the reference path untouched)."""
    import domain_specific_image_compression_b200 as sic
    from domain_specific_image_compression_b200 import layers as L
    torch.manual_seed(6)
    m = sic.CompressionModel(N=32, M=48, spatial_params=False, min_nu=2.0, max_nu=100.0).cuda()
    with torch.no_grad():
        m.g_a.g_a[14].weight.mul_(40.0)
        m.h_a.h_a[6].weight.mul_(40.0)
    x = torch.rand(2, 3, 64, 80, device="cuda")
    ny, nz = torch.rand(2, 48, 4, 5, device="cuda") - 0.5, torch.rand(2, 32, 1, 2, device="cuda") - 0.5
    res = {}
    try:
        for fast in (False, True):
            L.FAST_LAST_LAYER = fast
            m.train()
            m.zero_grad(set_to_none=True)
            out = m(x, "noise", noise_y=ny, noise_z=nz)
            loss, _, _ = sic.rate_distortion_loss(out, x, 100.0, "mse")
            loss.backward()
            res[fast] = (float(loss), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}, out["x_hat"].clone())
        m.eval()
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
            e_on = m(x, "round")["x_hat"]
            L.FAST_LAST_LAYER = False
            e_off = m(x, "round")["x_hat"]
        assert torch.equal(e_on, e_off)
    finally:
        L.FAST_LAST_LAYER = False
    (l0, g0, x0), (l1, g1, x1) = res[False], res[True]
    assert abs(l0 - l1) <= 1e-5 * abs(l0)
    assert float((x0 - x1).abs().max()) <= 1e-5 * float(x0.abs().max())
    assert g0.keys() == g1.keys()
    for n in g0:
        assert float((g0[n] - g1[n]).abs().max()) <= 1e-4 * float(g0[n].abs().max()) + 1e-8, n
