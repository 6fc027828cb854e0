"""The drop-in model API on the GPU: forward()/loss against the reference's golden run, compress()/decompress() round trips."""
import numpy as np
import pytest
import torch

from oracle import clib
from oracle import numpy_ref as R
from oracle import torch_port as TP

pytestmark = pytest.mark.gpu


def _model(golden, **kw):
    import domain_specific_image_compression_b200 as sic
    G = golden("model_small")
    m = sic.CompressionModel(N=16, M=24, spatial_params=False, min_nu=2.0, max_nu=100.0, **kw).cuda()
    m.load_state_dict({k[3:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("sd.")}, strict=True)
    return m, G


@pytest.fixture(autouse=True)
def _fp32_convs():
    """Both sides of a parity test pin the same conv arithmetic: full fp32, no TF32."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_forward_eval_vs_reference_golden(golden):
    import domain_specific_image_compression_b200 as sic
    m, G = _model(golden)
    m.eval()
    x = torch.from_numpy(G["x"]).cuda()
    with torch.no_grad():
        out = m(x, quant_mode="round")
        loss, Rr, D = sic.rate_distortion_loss(out, x, lambda_rd=100.0, dist="mse")
    assert set(out.keys()) == {"x_hat", "nll_y", "nll_z", "y", "y_tilde", "z", "z_tilde", "sigma", "nu"}
    assert out["sigma"].shape == out["y"].shape and out["sigma"].stride()[2:] == (0, 0)      # stride-0 expanded views, as the reference
    # convolutions run on cuDNN (GPU) vs MKL-DNN (golden, CPU): latents agree to conv rounding, not bit for bit
    np.testing.assert_allclose(out["y"].cpu().numpy(), G["eval.y"], rtol=1e-3, atol=2e-3)
    frac_equal = (out["y_tilde"].cpu().numpy() == G["eval.y_tilde"]).mean()
    assert frac_equal > 0.995
    np.testing.assert_allclose(out["sigma"].cpu().numpy(), G["eval.sigma"], rtol=2e-2)
    assert abs(float(Rr) - float(G["eval.R"])) < 0.02 * float(G["eval.R"])
    # the per-patch bit counts attached to the nll maps are the sums of those maps (regression: a GDN backward that ran
    # earlier on the same stream must not disturb the retire counter of the fused rate reduction)
    np.testing.assert_allclose(out["nll_y"]._sic_bits.cpu().numpy(), out["nll_y"].double().sum(dim=(1, 2, 3)).cpu().numpy(), rtol=2e-6)
    assert abs(float(D) - float(G["eval.D"])) < 1e-3


def test_forward_matches_eager_port_on_same_gpu(golden):
    """Same weights, same input, same device, same cuDNN: our forward vs the reference's op chains in eager PyTorch.
    Quantised latents bit-exact (north_star), nll within tolerance, bpp within 1e-5, x_hat bit-exact."""
    import domain_specific_image_compression_b200 as sic
    m, G = _model(golden)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    x = torch.from_numpy(G["x"]).cuda()
    with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True), torch.no_grad():
        m.eval()
        out = m(x, "round")
        ref = TP.forward(sd, x, "round", training=False)
        for k in ("y", "z", "y_tilde", "z_tilde", "x_hat"):
            assert torch.equal(out[k], ref[k]), k
        # P1 / N4: sigma, nu come from the fused hyper-synthesis tail (fixed summation order; a library GEMM cannot be matched bit for
        # bit).  Measured on B200: the eager chain's tiny 1x1 convolutions on [B,N,1,1] are off by up to 3.6e-5 relative from float64
        # (and from the reference's own CPU run in the golden file) whereas the kernel agrees with both to 2e-6 - so the tight bar is
        # the reference's golden sigma/nu, the loose one the eager GPU chain; with the tail switched back to the eager chain the
        # outputs are bit-equal to it
        for k in ("sigma", "nu"):
            assert out[k].shape == ref[k].shape and out[k].stride() == ref[k].stride()
            assert float(((out[k] - ref[k]).abs() / ref[k].abs()).max()) <= 1e-4, k
        # (latents on the GPU differ from the CPU golden run by conv rounding, which moves z_tilde and hence sigma: 2e-2 as above)
        np.testing.assert_allclose(out["sigma"].cpu().numpy(), G["eval.sigma"], rtol=2e-2)
        from domain_specific_image_compression_b200 import model as M_
        M_.FUSE_HYPER_TAIL = False
        try:
            out_e = m(x, "round")
        finally:
            M_.FUSE_HYPER_TAIL = True
        for k in ("sigma", "nu", "y_tilde", "x_hat"):
            assert torch.equal(out_e[k], ref[k]), k
        for k in ("nll_y", "nll_z"):
            err = (out[k].double() - ref[k].double()).abs()
            assert bool((err <= 1e-4 + 1e-5 * ref[k].double().abs()).all()), k
        l1, R1, D1 = sic.rate_distortion_loss(out, x, 100.0, "mse")
        l2, R2, D2 = TP.loss_fn(ref, x, 100.0, "mse")
        assert abs(float(R1) - float(R2)) <= 1e-5 * float(R2) and float(D1) == float(D2)
        m.train()
        ny, nz = torch.from_numpy(G["train.noise_y"]).cuda(), torch.from_numpy(G["train.noise_z"]).cuda()
        out = m(x, "noise", noise_y=ny, noise_z=nz)
        ref = TP.forward(sd, x, "noise", training=True, noise_y=ny, noise_z=nz)
        for k in ("y_tilde", "z_tilde", "x_hat"):
            assert torch.equal(out[k], ref[k]), k
        l1, R1, D1 = sic.rate_distortion_loss(out, x, 100.0, "msssim")
        l2, R2, D2 = TP.loss_fn(ref, x, 100.0, "msssim")          # the oracle's own MS-SSIM restatement
        assert abs(float(R1) - float(R2)) <= 1e-5 * float(R2) and abs(float(D1) - float(D2)) <= 1e-4


def test_training_gradients_vs_eager_port(golden):
    import domain_specific_image_compression_b200 as sic
    m, G = _model(golden)
    m.train()
    x = torch.from_numpy(G["x"]).cuda()
    ny, nz = torch.from_numpy(G["train.noise_y"]).cuda(), torch.from_numpy(G["train.noise_z"]).cuda()
    with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
        out = m(x, "noise", noise_y=ny, noise_z=nz)
        loss, _, _ = sic.rate_distortion_loss(out, x, 100.0, "mse")
        loss.backward()
        sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()}
        ref = TP.forward(sd, x, "noise", training=True, noise_y=ny, noise_z=nz)
        l2, _, _ = TP.loss_fn(ref, x, 100.0, "mse")
        l2.backward()
    assert abs(float(loss) - float(l2)) <= 1e-5 * abs(float(l2))
    checked = 0
    for name, p in m.named_parameters():
        if name.endswith(".gamma"):
            assert p.grad is None and sd[name].grad is None          # dead CxC parameter
            continue
        g_ref = sd[name].grad
        scale = float(g_ref.abs().max()) + 1e-12
        assert float((p.grad - g_ref).abs().max()) <= 2e-3 * scale, name
        checked += 1
    assert checked == 90 - 13
    # and against the reference's own CPU autograd (conv rounding differs: loose)
    for name, p in m.named_parameters():
        if "grad." + name in G.files:
            g_ref = torch.from_numpy(G["grad." + name]).cuda()
            assert float((p.grad - g_ref).abs().max()) <= 5e-2 * (float(g_ref.abs().max()) + 1e-9), name


@pytest.mark.parametrize("spatial", [False, True])
def test_compress_decompress_round_trip(golden, spatial):
    """cfg3 in miniature: encode -> bytes -> decode gives back exactly the quantised latents; x_hat == forward's x_hat;
    the tables/symbols the bytes were coded with equal the oracle's; the oracle coder produces the same bytes."""
    import domain_specific_image_compression_b200 as sic
    from domain_specific_image_compression_b200 import functional as F
    if spatial:
        torch.manual_seed(0)
        m = sic.CompressionModel(N=16, M=24, spatial_params=True, min_nu=2.0).cuda()
        with torch.no_grad():
            m.g_a.g_a[14].weight.mul_(40.0); m.h_a.h_a[6].weight.mul_(40.0)
        G = golden("model_small")
    else:
        m, G = _model(golden)
    m.eval()
    x = torch.nn.functional.interpolate(torch.rand(3, 3, 16, 16, generator=torch.Generator().manual_seed(5)), size=(128, 128), mode="bilinear").clamp(0, 1).cuda()
    comp = m.compress(x, tail=10)
    assert set(comp.keys()) == {"strings", "shape_y", "shape_z", "min_y", "max_y", "min_z", "max_z"}
    assert len(comp["strings"]) == 3 and all(len(s) == 2 and isinstance(s[0], bytes) for s in comp["strings"])
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
        out = m(x, "round")
    x_hat = m.decompress(comp)
    assert torch.equal(x_hat, out["x_hat"].clamp(0, 1))                               # decoded reconstruction == forward's
    # oracle cross-check of one patch: symbols, tables, bytes
    b = 1
    yq, zq = out["y_tilde"].cpu().numpy(), out["z_tilde"].cpu().numpy()
    sy, mny, mxy = R.symbols_and_support(yq, 10)
    sz, mnz, mxz = R.symbols_and_support(zq, 10)
    assert comp["min_y"] == mny.tolist() and comp["max_y"] == mxy.tolist() and comp["min_z"] == mnz.tolist() and comp["max_z"] == mxz.tolist()
    C, h, w = yq.shape[1:]
    if spatial:
        sig, nu = out["sigma"][b].cpu().numpy().ravel(), out["nu"][b].cpu().numpy().ravel()
        spr = 1
    else:
        sig, nu = out["sigma"][b, :, 0, 0].cpu().numpy(), out["nu"][b, :, 0, 0].cpu().numpy()
        spr = h * w
    ty = clib.build_tables("studentt", sig, nu, np.zeros(sig.size, np.int32), mny[b:b + 1], mxy[b:b + 1])
    assert clib.rans_encode(sy[b], ty, int(mxy[b] - mny[b] + 1), spr) == comp["strings"][b][1]
    lsz = m.z_prior.log_sigma.detach().cpu().numpy()
    tz = clib.build_tables("gaussian", clib.exp_f32(lsz), None, np.zeros(lsz.size, np.int32), mnz[b:b + 1], mxz[b:b + 1])
    assert clib.rans_encode(sz[b], tz, int(mxz[b] - mnz[b] + 1), zq.shape[2] * zq.shape[3]) == comp["strings"][b][0]
    # real rate vs estimated rate: coded bits within a few % + header of the discretised-model cross entropy is not required,
    # but the stream must be shorter than raw int16 symbols
    nbytes = sum(len(s) for pair in comp["strings"] for s in pair)
    assert nbytes < 2 * (yq.size + zq.size)


def test_stream_erasure_is_detected(golden):
    import domain_specific_image_compression_b200 as sic
    m, _ = _model(golden)
    x = torch.rand(1, 3, 128, 128, generator=torch.Generator().manual_seed(1)).cuda()
    comp = m.compress(x)
    comp["strings"][0][1] = comp["strings"][0][1][:-6]
    with pytest.raises(sic.SicError):
        m.decompress(comp)


@pytest.mark.parametrize("N", [128, 192])
def test_model_with_dense_gamma_gdn_on_tensor_cores(N):
    """north_star's G3 inside the model: every GDN/IGDN site switched to the dense C x C gamma (tcgen05 kernel, C = N).  With the
    reference's initialisation (layers.py:13) gamma is diagonal, so the dense model must reproduce the reference-path model
    up to gamma at TF32 precision; gradients flow to the C x C matrices instead of the depthwise weights."""
    import domain_specific_image_compression_b200 as sic
    torch.manual_seed(N)
    m = sic.CompressionModel(N=N, M=64, spatial_params=False, min_nu=2.0, max_nu=100.0).cuda()
    with torch.no_grad():
        m.g_a.g_a[14].weight.mul_(20.0)             # spread the latents (default init gives y ~ 0: SURVEY 8(d))
    x = torch.rand(2, 3, 64, 64, device="cuda")
    noise_y = torch.rand(2, 64, 4, 4, device="cuda") - 0.5
    noise_z = torch.rand(2, N, 1, 1, device="cuda") - 0.5
    ref = m(x, "noise", noise_y=noise_y, noise_z=noise_z)
    sites = [mod for mod in m.modules() if isinstance(mod, sic.GDN)]
    assert len(sites) == 13
    for mod in sites:
        mod.dense = True
    out = m(x, "noise", noise_y=noise_y, noise_z=noise_z)
    for k in ("y", "x_hat"):
        a, b = out[k], ref[k]
        assert float((a - b).abs().max()) <= 5e-3 * float(b.abs().max()) + 1e-6, k
    loss, _, _ = sic.rate_distortion_loss(out, x, lambda_rd=100.0, dist="mse")
    loss.backward()
    for mod in sites:
        assert mod.gamma.grad is not None and torch.isfinite(mod.gamma.grad).all() and float(mod.gamma.grad.abs().max()) > 0
        assert mod.gamma_conv.weight.grad is None


def test_cfg3_full_size_round_trip_and_erasure():
    """BASELINE.json configs[2] at its own size: N=128, M=192 model on 512x512 patches.  encode -> decode reproduces forward()'s
    reconstruction bit for bit (both coders emit the same bytes), and dropping the tail of any stream is detected."""
    import domain_specific_image_compression_b200 as sic
    torch.manual_seed(3)
    m = sic.CompressionModel(N=128, M=192, spatial_params=False, min_nu=2.0, max_nu=100.0).cuda().eval()
    with torch.no_grad():
        m.g_a.g_a[14].weight.mul_(40.0)             # spread latents: default init gives y ~ 0
        m.h_a.h_a[6].weight.mul_(40.0)
    x = torch.nn.functional.interpolate(torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(9)), size=(512, 512),
                                        mode="bilinear").clamp(0, 1).cuda()
    comp = m.compress(x, tail=10)
    assert comp["shape_y"] == [2, 192, 32, 32] and comp["shape_z"] == [2, 128, 8, 8]
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
        want = m(x, "round")["x_hat"].clamp(0, 1)
    from domain_specific_image_compression_b200 import container
    blob = container.pack(comp)                                   # SIC-CONT-1: the dict as one self-describing byte string
    assert container.unpack(blob) == comp
    assert torch.equal(m.decompress(container.unpack(blob)), want)
    assert m.compress(x, tail=10, coder="host")["strings"] == comp["strings"]
    bad = dict(comp)
    bad["strings"] = [[s[0], s[1][: len(s[1]) // 2]] for s in comp["strings"]]
    with pytest.raises(sic.SicError):
        m.decompress(bad)


def test_training_step_under_autocast_like_the_reference(golden):
    """train.py:196-204 runs forward + loss under torch.cuda.amp.autocast with a GradScaler (config.py TRAIN.amp=True).  The cuDNN
    convs then return float16; our kernels take it (cast up, computed in float32) instead of raising, and every live
    parameter receives a finite float32 gradient.  Loss within float16 conv error of the float32 run."""
    import domain_specific_image_compression_b200 as sic
    m, G = _model(golden)
    m.train()
    x = torch.from_numpy(G["x"]).cuda()
    ny, nz = torch.from_numpy(G["train.noise_y"]).cuda(), torch.from_numpy(G["train.noise_z"]).cuda()
    out32 = m(x, quant_mode="noise", noise_y=ny, noise_z=nz)
    loss32, _, _ = sic.rate_distortion_loss(out32, x, lambda_rd=100.0, dist="msssim")
    scaler = torch.amp.GradScaler("cuda", init_scale=8.0)      # the default 65536 overflows float16 on the first steps by design
    with torch.autocast("cuda", dtype=torch.float16):
        out = m(x, quant_mode="noise", noise_y=ny, noise_z=nz)
        loss, Rr, D = sic.rate_distortion_loss(out, x, lambda_rd=100.0, dist="msssim")
    assert out["y_tilde"].dtype == torch.float32 and out["nll_y"].dtype == torch.float32     # kernel outputs stay float32
    assert abs(float(loss) - float(loss32)) < 0.05 * abs(float(loss32)) + 1e-2
    scaler.scale(loss).backward()
    live = [(n, p) for n, p in m.named_parameters() if not n.endswith(".gamma")]
    assert all(p.grad is not None and p.grad.dtype == torch.float32 and torch.isfinite(p.grad).all() for _, p in live)


def test_overlapped_hyper_branch_matches_sequential_schedule(golden):
    """model.OVERLAP_HYPER_BRANCH: the hyperprior branch on a side stream next to the synthesis transform.  With the same supplied
    noise the nine outputs are bit-identical to the sequential schedule (y_tilde = y + noise either way; the likelihood kernel sees
    the same y_tilde), the loss is equal and the gradients agree to summation order (three gradient contributions meet at y)."""
    import domain_specific_image_compression_b200 as sic
    from domain_specific_image_compression_b200 import model as M_
    m, G = _model(golden)
    m.train()
    x = torch.from_numpy(G["x"]).cuda()
    ny, nz = torch.from_numpy(G["train.noise_y"]).cuda(), torch.from_numpy(G["train.noise_z"]).cuda()
    res = {}
    try:
        for ov in (False, True):
            M_.OVERLAP_HYPER_BRANCH = ov
            m.zero_grad(set_to_none=True)
            with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
                out = m(x, "noise", noise_y=ny, noise_z=nz)
                loss, Rr, D = sic.rate_distortion_loss(out, x, 100.0, "msssim")
                loss.backward()
            torch.cuda.synchronize()
            res[ov] = ({k: v.detach().clone() for k, v in out.items()}, float(loss), float(Rr),
                       {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
        # Philox path (no supplied noise): runs, finite, fresh noise per call
        M_.OVERLAP_HYPER_BRANCH = True
        a = m(x, "noise")
        b = m(x, "noise")
        torch.cuda.synchronize()
        assert not torch.equal(a["y_tilde"], b["y_tilde"]) and torch.isfinite(a["nll_y"]).all()
        assert float((a["y_tilde"] - a["y"]).abs().max()) <= 0.5
    finally:
        M_.OVERLAP_HYPER_BRANCH = False
    (o0, l0, r0, g0), (o1, l1, r1, g1) = res[False], res[True]
    for k in o0:
        assert torch.equal(o0[k], o1[k]), k
    assert l0 == l1 and r0 == r1
    assert g0.keys() == g1.keys()
    for n in g0:
        assert float((g0[n] - g1[n]).abs().max()) <= 1e-5 * float(g0[n].abs().max()) + 1e-9, n
