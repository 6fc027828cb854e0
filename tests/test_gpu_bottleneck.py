"""K1 (fused quantise + likelihood + rate, and its backward) on the GPU through the C ABI, against the oracle and the
reference's golden outputs.  Tolerances: y_tilde bit-exact; nll per element |d| <= 1e-4 bits + 1e-5*|nll| (the reference's
own fp32-vs-fp64 error is 7e-5 bits); bits/bpp 1e-5 relative (north_star)."""
import numpy as np
import pytest
import torch

from oracle import numpy_ref as R

pytestmark = pytest.mark.gpu


def _F():
    from domain_specific_image_compression_b200 import functional as F
    return F


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def assert_nll_close(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    assert (err <= 1e-4 + 1e-5 * np.abs(ref)).all(), f"max err {err.max()}"


def test_density_broadcast_vs_reference_golden(golden):
    G = golden("likelihood")
    F = _F()
    x = dev(G["x"])
    yt, nll, bits = F.bottleneck(x, dev(G["sigma_bc"]), dev(G["nu_bc"]), quant="none", lik="density")
    assert yt.data_ptr() == x.data_ptr()
    assert_nll_close(nll.cpu().numpy(), G["nll_bc"])                                        # the reference itself
    assert_nll_close(nll.cpu().numpy(), R.studentt_nll_f64(G["x"], G["sigma_bc"], G["nu_bc"]))  # float64 truth
    ref_bits = G["nll_bc"].astype(np.float64).sum(axis=(1, 2, 3))
    np.testing.assert_allclose(bits.cpu().numpy(), ref_bits, rtol=1e-5)
    # bits is exactly the fixed-order sum of the nll the kernel wrote (to fp32 rounding of the fp64 fold)
    np.testing.assert_allclose(bits.cpu().numpy(), nll.double().sum(dim=(1, 2, 3)).cpu().numpy(), rtol=2e-6)


def test_density_spatial_vs_reference_golden(golden):
    G = golden("likelihood")
    F = _F()
    _, nll, bits = F.bottleneck(dev(G["x"]), dev(G["sigma_sp"]), dev(G["nu_sp"]), quant="none", lik="density")
    assert_nll_close(nll.cpu().numpy(), G["nll_sp"])
    np.testing.assert_allclose(bits.cpu().numpy(), G["nll_sp"].astype(np.float64).sum(axis=(1, 2, 3)), rtol=1e-5)


def test_gaussian_vs_reference_golden(golden):
    G = golden("likelihood")
    F = _F()
    _, nll, bits = F.bottleneck(dev(G["z"]), dev(G["log_sigma_z"]), quant="none", lik="gaussian")
    got, ref = nll.cpu().numpy().astype(np.float64), G["nll_z"].astype(np.float64)
    assert (np.abs(got - ref) <= 1e-4 + 1e-5 * np.abs(ref)).all()
    np.testing.assert_allclose(bits.cpu().numpy(), ref.sum(axis=(1, 2, 3)), rtol=1e-5)


def test_quantize_round_and_noise_bit_exact(golden):
    G = golden("likelihood")
    F = _F()
    y = G["x"].copy()
    y.ravel()[:6] = [0.5, 1.5, 2.5, -0.5, -0.4, -2.5]                    # half-to-even ties and -0.0
    yt, nll, _ = F.bottleneck(dev(y), dev(G["sigma_bc"]), dev(G["nu_bc"]), quant="round")
    ref = R.quantize_round(y)
    assert np.array_equal(yt.cpu().numpy().view(np.uint32), ref.view(np.uint32))            # incl. the sign of -0.0
    assert_nll_close(nll.cpu().numpy(), R.studentt_nll_f64(ref, G["sigma_bc"], G["nu_bc"]))
    noise = (np.random.default_rng(0).random(y.shape, dtype=np.float32) - np.float32(0.5))
    yt, nll, _ = F.bottleneck(dev(y), dev(G["sigma_bc"]), dev(G["nu_bc"]), quant="noise", noise=dev(noise))
    ref = R.quantize_noise(y, noise)
    assert np.array_equal(yt.cpu().numpy(), ref)
    assert_nll_close(nll.cpu().numpy(), R.studentt_nll_f64(ref, G["sigma_bc"], G["nu_bc"]))
    with pytest.raises(ValueError):
        F.bottleneck(dev(y), dev(G["sigma_bc"]), dev(G["nu_bc"]), quant="floor")


def test_model_latents_from_reference_forward(golden):
    """Feed the reference model's own y, z, sigma, nu and noise draws (model_small fixture) through K1."""
    G = golden("model_small")
    F = _F()
    for mode in ("eval", "train"):
        kw = dict(quant="round") if mode == "eval" else dict(quant="noise", noise=dev(G["train.noise_y"]))
        sig, nu = G[mode + ".sigma"][:, :, :1, :1], G[mode + ".nu"][:, :, :1, :1]
        yt, nll, bits = F.bottleneck(dev(G[mode + ".y"]), dev(sig), dev(nu), **kw)
        assert np.array_equal(yt.cpu().numpy(), G[mode + ".y_tilde"])                       # quantised latents bit-exact
        assert_nll_close(nll.cpu().numpy(), G[mode + ".nll_y"])
        kw = dict(quant="round") if mode == "eval" else dict(quant="noise", noise=dev(G["train.noise_z"]))
        zt, nz, bz = F.bottleneck(dev(G[mode + ".z"]), dev(G["sd.z_prior.log_sigma"]), lik="gaussian", **kw)
        assert np.array_equal(zt.cpu().numpy(), G[mode + ".z_tilde"])
        assert_nll_close(nz.cpu().numpy(), G[mode + ".nll_z"])
        x = G["x"]
        bpp = max(float(bits.double().sum() + bz.double().sum()) / (x.shape[0] * x.shape[2] * x.shape[3]), 0.0)
        assert abs(bpp - float(G[mode + ".R"])) <= 1e-5 * float(G[mode + ".R"])              # R1: bpp within 1e-5 rel


def test_backward_vs_reference_autograd(golden):
    G = golden("likelihood")
    F = _F()
    x = dev(G["x"]).requires_grad_(True)
    sig = dev(G["sigma_bc"]).requires_grad_(True)
    nu = dev(G["nu_bc"]).requires_grad_(True)
    _, nll, _ = F.bottleneck(x, sig, nu, quant="none")
    (nll * dev(G["g"])).sum().backward()
    dx, ds, dn = R.studentt_nll_grads_f64(G["x"], G["sigma_bc"], G["nu_bc"], G["g"])
    scale = np.abs(dx).max()
    np.testing.assert_allclose(x.grad.cpu().numpy(), dx, rtol=2e-5, atol=2e-6 * scale)
    np.testing.assert_allclose(x.grad.cpu().numpy(), G["dx_bc"], rtol=1e-4, atol=1e-5 * scale)       # the reference's autograd
    ds_ref, dn_ref = ds.sum((2, 3), keepdims=True), dn.sum((2, 3), keepdims=True)
    np.testing.assert_allclose(sig.grad.cpu().numpy(), ds_ref, rtol=1e-4, atol=1e-5 * np.abs(ds_ref).max())
    np.testing.assert_allclose(nu.grad.cpu().numpy(), dn_ref, rtol=1e-3, atol=1e-5)
    # clamp masks (closed interval): fixture entries 0/3 are outside, 1/2 exactly on the bounds
    assert sig.grad.view(-1)[0] == 0 and sig.grad.view(-1)[3] == 0 and sig.grad.view(-1)[1] != 0 and sig.grad.view(-1)[2] != 0
    assert nu.grad.view(-1)[0] == 0 and nu.grad.view(-1)[3] == 0 and nu.grad.view(-1)[1] != 0 and nu.grad.view(-1)[2] != 0
    # spatial layout
    x.grad = None
    sig = dev(G["sigma_sp"]).requires_grad_(True)
    nu = dev(G["nu_sp"]).requires_grad_(True)
    _, nll, _ = F.bottleneck(x, sig, nu, quant="none")
    (nll * dev(G["g"])).sum().backward()
    dx, ds, dn = R.studentt_nll_grads_f64(G["x"], G["sigma_sp"], G["nu_sp"], G["g"])
    np.testing.assert_allclose(x.grad.cpu().numpy(), dx, rtol=2e-5, atol=2e-6 * np.abs(dx).max())
    np.testing.assert_allclose(sig.grad.cpu().numpy(), ds, rtol=1e-4, atol=1e-5 * np.abs(ds).max())
    np.testing.assert_allclose(nu.grad.cpu().numpy(), dn, rtol=1e-3, atol=2e-6)


def test_backward_gaussian_bits_and_round(golden):
    G = golden("likelihood")
    F = _F()
    z = dev(G["z"]).requires_grad_(True)
    ls = dev(G["log_sigma_z"]).requires_grad_(True)
    _, nll, _ = F.bottleneck(z, ls, quant="none", lik="gaussian")
    (nll * dev(G["gz"])).sum().backward()
    dz, dls = R.gaussian_nll_grads_f64(G["z"], G["log_sigma_z"], G["gz"])
    np.testing.assert_allclose(z.grad.cpu().numpy(), dz, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(ls.grad.cpu().numpy(), dls, rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(ls.grad.cpu().numpy(), G["dlog_sigma_z"], rtol=1e-4, atol=1e-3)
    assert ls.grad[0] == 0 and ls.grad[1] == 0
    # upstream through the per-patch bit counts (what rate_distortion_loss uses) == upstream through nll.sum()
    x = dev(G["x"]).requires_grad_(True)
    sig = dev(G["sigma_bc"]).requires_grad_(True)
    nu = dev(G["nu_bc"]).requires_grad_(True)
    w = torch.tensor([1.0, -2.0, 0.5, 3.0], device="cuda")
    yt, nll, bits = F.bottleneck(x, sig, nu, quant="noise", noise=torch.zeros_like(x))
    ((bits * w).sum() + (yt * 0.25).sum()).backward()
    gx, gs, gn = x.grad.clone(), sig.grad.clone(), nu.grad.clone()
    dx, ds, dn = R.studentt_nll_grads_f64(G["x"], G["sigma_bc"], G["nu_bc"], w.cpu().numpy().reshape(4, 1, 1, 1))
    np.testing.assert_allclose(gx.cpu().numpy(), dx + 0.25, rtol=2e-5, atol=2e-6 * np.abs(dx).max())
    np.testing.assert_allclose(gs.cpu().numpy(), ds.sum((2, 3), keepdims=True), rtol=1e-4, atol=1e-5 * np.abs(ds).max())
    # round: torch.round has zero gradient, no straight-through (model.py:32-33)
    x.grad = None
    yt, nll, bits = F.bottleneck(x, sig, nu, quant="round")
    (bits.sum() + yt.sum()).backward()
    assert float(x.grad.abs().max()) == 0.0


def test_philox_noise_matches_oracle_and_advances():
    F = _F()
    torch.manual_seed(1234)
    y = torch.zeros(2, 3, 5, 4, device="cuda")
    sig = torch.ones(2, 3, 1, 1, device="cuda")
    nu = torch.full((2, 3, 1, 1), 4.0, device="cuda")
    st = F.philox_state(y.device)
    assert st.tolist() == [1234, 0]
    n1, _, _ = F.bottleneck(y, sig, nu, quant="noise")
    assert st.tolist() == [1234, 1]                                       # the kernel advanced the offset itself
    n2, _, _ = F.bottleneck(y, sig, nu, quant="noise")
    assert np.array_equal(n1.cpu().numpy().ravel(), R.philox_uniform_noise(y.numel(), 1234, 0))
    assert np.array_equal(n2.cpu().numpy().ravel(), R.philox_uniform_noise(y.numel(), 1234, 1))
    assert float(n1.abs().max()) < 0.5
    # odd HW takes the scalar path: same stream of noise
    y = torch.zeros(1, 2, 3, 3, device="cuda")
    torch.manual_seed(7)
    n3, _, _ = F.bottleneck(y, torch.ones(1, 2, 1, 1, device="cuda"), torch.full((1, 2, 1, 1), 4.0, device="cuda"), quant="noise")
    assert np.array_equal(n3.cpu().numpy().ravel(), R.philox_uniform_noise(18, 7, 0))
    big = torch.zeros(4, 64, 64, 64, device="cuda")
    nb, _, _ = F.bottleneck(big, torch.ones(4, 64, 1, 1, device="cuda"), torch.full((4, 64, 1, 1), 4.0, device="cuda"), quant="noise")
    assert abs(float(nb.mean())) < 2e-3 and abs(float(nb.var()) - 1 / 12) < 2e-3


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (2, 3, 1, 1), (1, 5, 3, 3), (3, 7, 5, 9), (2, 4, 2, 2), (1, 2, 130, 129)])
def test_ragged_shapes(shape):
    """HW not a multiple of 4 (scalar path), HW = 1 (64x64 images give 1x1 hyper-latents), segments that straddle."""
    F = _F()
    rng = np.random.default_rng(sum(shape))
    B, C = shape[:2]
    y = (rng.standard_normal(shape) * 3).astype(np.float32)
    sig = np.exp(rng.standard_normal((B, C, 1, 1))).astype(np.float32)
    nu = rng.uniform(2, 50, (B, C, 1, 1)).astype(np.float32)
    yt, nll, bits = F.bottleneck(dev(y), dev(sig), dev(nu), quant="round")
    ref = R.studentt_nll_f64(R.quantize_round(y), sig, nu)
    assert_nll_close(nll.cpu().numpy(), ref)
    np.testing.assert_allclose(bits.cpu().numpy(), ref.sum(axis=(1, 2, 3)), rtol=1e-5, atol=1e-4)
    ls = rng.standard_normal(C).astype(np.float32)
    zt, nz, bz = F.bottleneck(dev(y), dev(ls), quant="round", lik="gaussian")
    refz = R.gaussian_nll_f64(R.quantize_round(y), ls)
    assert (np.abs(nz.cpu().numpy() - refz) <= 1e-4 + 1e-5 * np.abs(refz)).all()
    np.testing.assert_allclose(bz.cpu().numpy(), refz.sum(axis=(1, 2, 3)), rtol=1e-5, atol=1e-4)


def test_location_parameter():
    """north_star's mu (the reference has no location head: SURVEY D2) — shift invariance against the oracle."""
    F = _F()
    rng = np.random.default_rng(5)
    y = (rng.standard_normal((2, 6, 8, 8)) * 3).astype(np.float32)
    sig = np.exp(rng.standard_normal((2, 6, 1, 1))).astype(np.float32)
    nu = rng.uniform(2, 50, (2, 6, 1, 1)).astype(np.float32)
    mu = rng.standard_normal((2, 6, 1, 1)).astype(np.float32)
    ydev, mudev = dev(y).requires_grad_(True), dev(mu).requires_grad_(True)
    _, nll, _ = F.bottleneck(ydev, dev(sig), dev(nu), mudev, quant="none")
    assert_nll_close(nll.detach().cpu().numpy(), R.studentt_nll_f64(y - mu, sig, nu))
    nll.sum().backward()
    np.testing.assert_allclose(mudev.grad.cpu().numpy(), -ydev.grad.sum(dim=(2, 3), keepdim=True).cpu().numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("B,C,h,w", [(16, 192, 16, 16), (64, 320, 16, 16), (1, 320, 128, 128)])
def test_full_size_properties(B, C, h, w):
    """BASELINE.json sizes (cfg2 / cfg4 / top of the cfg5 sweep): size-independent properties instead of the slow oracle:
    bits == sum(nll) per patch; rounding is idempotent; permuting patches permutes bits; a sampled slab matches the oracle."""
    F = _F()
    g = torch.Generator(device="cuda").manual_seed(0)
    y = torch.randn(B, C, h, w, device="cuda", generator=g) * 3
    sig = torch.exp(torch.randn(B, C, 1, 1, device="cuda", generator=g))
    nu = torch.exp(torch.randn(B, C, 1, 1, device="cuda", generator=g) + 1.5)
    yt, nll, bits = F.bottleneck(y, sig, nu, quant="round")
    np.testing.assert_allclose(bits.cpu().numpy(), nll.double().sum(dim=(1, 2, 3)).cpu().numpy(), rtol=2e-6)
    yt2, nll2, bits2 = F.bottleneck(yt, sig, nu, quant="round")
    assert torch.equal(yt2, yt) and torch.equal(nll2, nll) and torch.equal(bits2, bits)      # idempotent + deterministic
    if B > 1:
        perm = torch.arange(B - 1, -1, -1, device="cuda")
        _, _, bits_p = F.bottleneck(y[perm].contiguous(), sig[perm].contiguous(), nu[perm].contiguous(), quant="round")
        assert torch.equal(bits_p, bits[perm])
    sl = (slice(0, 1), slice(0, 8))
    ref = R.studentt_nll_f64(yt[sl].cpu().numpy(), sig[sl].cpu().numpy(), nu[sl].cpu().numpy())
    assert_nll_close(nll[sl].cpu().numpy(), ref)


# ------------------------------------------------------------------------------------------------ L2: cdf_diff likelihood
def _cdf_case(rng, shape, spatial):
    B, C = shape[:2]
    pshape = shape if spatial else (B, C, 1, 1)
    sig = np.exp(rng.uniform(np.log(2e-2), np.log(50.0), pshape)).astype(np.float32)
    nu = np.exp(rng.uniform(np.log(2.0), np.log(100.0), pshape)).astype(np.float32)
    y = (rng.standard_t(3, shape) * sig * rng.choice([1.0, 4.0], shape)).astype(np.float32)
    return y, sig, nu


@pytest.mark.parametrize("spatial", [False, True])
def test_cdfdiff_forward_vs_scipy(spatial):
    """P = T_nu(y+1/2) - T_nu(y-1/2) (north_star); oracle: scipy float64 same-side survival differences."""
    F = _F()
    rng = np.random.default_rng(11 + spatial)
    y, sig, nu = _cdf_case(rng, (3, 12, 8, 8), spatial)
    for quant, yq in (("none", y), ("round", R.quantize_round(y))):
        yt, nll, bits = F.bottleneck(dev(y), dev(sig), dev(nu), quant=quant, lik="cdf_diff")
        ref = R.studentt_cdfdiff_nll_f64(yq, sig, nu)
        assert_nll_close(nll.cpu().numpy(), ref)
        np.testing.assert_allclose(bits.cpu().numpy(), ref.sum(axis=(1, 2, 3)), rtol=1e-5)
    # probabilities of the integer support sum to one (the table property the coder relies on)
    k = np.arange(-3000, 3001, dtype=np.float32).reshape(1, 1, -1, 1)
    _, nll, _ = F.bottleneck(dev(np.broadcast_to(k, (1, 3, 6001, 1)).copy()), dev(np.array([0.3, 1.0, 7.0], np.float32).reshape(1, 3, 1, 1)),
                             dev(np.array([2.0, 5.0, 40.0], np.float32).reshape(1, 3, 1, 1)), quant="none", lik="cdf_diff")
    total = torch.exp2(-nll.double()).sum(dim=2).view(-1).cpu().numpy()
    # nu=2, sigma=0.3: the mass beyond +-3000 is (sigma/3000)^2 ~ 1e-8; fp32 accumulation of 6001 terms dominates
    assert np.abs(total - 1.0).max() < 1e-5, total


def test_cdfdiff_extremes_and_clamps():
    """Far tails (no underflow / cancellation), sigma and nu outside the clamp range, wide bins (sigma << 1)."""
    F = _F()
    y = np.array([0.0, 0.4, -0.5, 3.0, 40.0, -250.0, 1e4, 0.0, 2.0, 5.0, -7.0, 300.0], np.float32).reshape(1, 12, 1, 1)
    sig = np.array([1e-4, 1e-3, 0.05, 0.3, 1.0, 2.0, 5.0, 2e3, 1e3, 0.7, 12.0, 0.9], np.float32).reshape(1, 12, 1, 1)
    nu = np.array([1.5, 2.0, 3.0, 100.0, 150.0, 2.5, 2.0, 7.0, 30.0, 80.0, 2.0, 60.0], np.float32).reshape(1, 12, 1, 1)
    _, nll, _ = F.bottleneck(dev(y), dev(sig), dev(nu), quant="none", lik="cdf_diff")
    ref = R.studentt_cdfdiff_nll_f64(y, sig, nu)
    got = nll.cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    assert (np.abs(got - ref) <= 2e-4 + 2e-5 * np.abs(ref)).all(), (got.ravel(), ref.ravel())


@pytest.mark.parametrize("spatial", [False, True])
def test_cdfdiff_backward_vs_float64_finite_differences(spatial):
    F = _F()
    rng = np.random.default_rng(21 + spatial)
    shape = (2, 6, 4, 4)
    y, sig, nu = _cdf_case(rng, shape, spatial)
    nu = np.clip(nu, 2.05, 95.0).astype(np.float32)                      # keep away from the clamp kinks for the FD reference
    g = rng.standard_normal(shape).astype(np.float32)
    yd, sd, nd = dev(y).requires_grad_(True), dev(sig).requires_grad_(True), dev(nu).requires_grad_(True)
    _, nll, _ = F.bottleneck(yd, sd, nd, quant="none", lik="cdf_diff")
    (nll * dev(g)).sum().backward()
    f = lambda yy, ss, nn: R.studentt_cdfdiff_nll_f64(yy, ss, nn)
    y64, s64, n64 = y.astype(np.float64), sig.astype(np.float64), nu.astype(np.float64)
    e = 1e-5
    dx = g * (f(y64 + e, s64, n64) - f(y64 - e, s64, n64)) / (2 * e)
    ds = g * (f(y64, s64 * (1 + e), n64) - f(y64, s64 * (1 - e), n64)) / (2 * e * s64)
    dn = g * (f(y64, s64, n64 * (1 + e)) - f(y64, s64, n64 * (1 - e))) / (2 * e * n64)
    if not spatial:
        ds, dn = ds.sum((2, 3), keepdims=True), dn.sum((2, 3), keepdims=True)
    np.testing.assert_allclose(yd.grad.cpu().numpy(), dx, rtol=2e-3, atol=2e-4 * np.abs(dx).max())
    np.testing.assert_allclose(sd.grad.cpu().numpy(), ds, rtol=2e-3, atol=2e-4 * np.abs(ds).max())
    np.testing.assert_allclose(nd.grad.cpu().numpy(), dn, rtol=5e-3, atol=5e-4 * np.abs(dn).max())


def test_cdfdiff_model_mode_trains(golden):
    """CompressionModel(likelihood='cdf_diff'): forward/backward runs, rate is >= 0 by construction (a probability mass,
    unlike the density rate which can go negative: SURVEY D1) and close to the density rate when sigma is not small."""
    import domain_specific_image_compression_b200 as sic
    G = golden("model_small")
    sd = {k[3:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("sd.")}
    x = torch.from_numpy(G["x"]).cuda()
    rates = {}
    for lik in ("density", "cdf_diff"):
        m = sic.CompressionModel(N=16, M=24, min_nu=2.0, likelihood=lik).cuda()
        m.load_state_dict(sd)
        m.train()
        out = m(x, "noise", noise_y=torch.from_numpy(G["train.noise_y"]).cuda(), noise_z=torch.from_numpy(G["train.noise_z"]).cuda())
        loss, Rr, D = sic.rate_distortion_loss(out, x, 100.0, "mse")
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for n, p in m.named_parameters() if p.grad is not None)
        rates[lik] = float(Rr)
        if lik == "cdf_diff":
            assert float(out["nll_y"].min()) >= 0.0
    assert abs(rates["cdf_diff"] - rates["density"]) < 0.25 * rates["density"]


def test_empty_inputs():
    """Empty batch / empty latent: the reference's op chains return empty tensors; so do we (no launch)."""
    F = _F()
    y = torch.zeros(0, 8, 4, 4, device="cuda")
    yt, nll, bits = F.bottleneck(y, torch.ones(0, 8, 1, 1, device="cuda"), torch.ones(0, 8, 1, 1, device="cuda"), quant="round")
    assert yt.shape == y.shape and nll.shape == y.shape and bits.shape == (0,)
    y = torch.zeros(2, 8, 0, 4, device="cuda")
    yt, nll, bits = F.bottleneck(y, torch.ones(2, 8, 1, 1, device="cuda"), torch.full((2, 8, 1, 1), 3.0, device="cuda"), quant="noise")
    assert nll.numel() == 0 and bits.tolist() == [0.0, 0.0]
    x = torch.zeros(0, 16, 8, 8, device="cuda")
    assert F.gdn(x, torch.ones(16, device="cuda"), torch.ones(16, 1, 1, 1, device="cuda")).shape == x.shape


def test_static_quantize_helper_runs_the_kernel():
    """CompressionModel.quantize (model.py:27-35) as a standalone call: half-to-even with the sign of zero kept, zero gradient
    for 'round' (torch.round), unit gradient and |noise| <= 1/2 for 'noise'."""
    import domain_specific_image_compression_b200 as sic
    x = torch.tensor([0.5, 1.5, 2.5, -0.5, -0.4, 3.49], device="cuda")
    q = sic.CompressionModel.quantize(x, "round")
    assert torch.equal(q.view(torch.int32), torch.tensor([0.0, 2.0, 2.0, -0.0, -0.0, 3.0], device="cuda").view(torch.int32))
    big = torch.randn(3, 5, 7, 11, device="cuda", requires_grad=True)
    n = sic.CompressionModel.quantize(big, "noise")
    assert n.shape == big.shape and float((n - big).abs().max()) <= 0.5 and float((n - big).abs().max()) > 0.3
    n.sum().backward()
    assert torch.equal(big.grad, torch.ones_like(big))
    big.grad = None
    sic.CompressionModel.quantize(big, "round").sum().backward()
    assert float(big.grad.abs().max()) == 0.0
