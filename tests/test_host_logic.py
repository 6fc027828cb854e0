"""Host-side logic that needs no GPU: API surface, state_dict layout, host entropy coder, MS-SSIM restatement."""
import numpy as np
import pytest
import torch
from scipy import ndimage

import domain_specific_image_compression_b200 as sic
from domain_specific_image_compression_b200 import functional as F
from domain_specific_image_compression_b200.losses import multi_scale_ssim
from oracle import clib


def test_state_dict_layout_matches_reference(golden):
    G = golden("model_small")
    ref_keys = sorted(k[3:] for k in G.files if k.startswith("sd."))
    m = sic.CompressionModel(N=16, M=24, min_nu=2.0)
    assert sorted(m.state_dict().keys()) == ref_keys and len(ref_keys) == 90
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == G["sd." + k].shape, k
    m.load_state_dict({k: torch.from_numpy(G["sd." + k]) for k in ref_keys}, strict=True)
    big = sic.CompressionModel()
    assert sum(p.numel() for p in big.parameters()) == 6_483_267
    sp = sic.CompressionModel(N=8, M=8, spatial_params=True).state_dict()
    assert "h_s.to_sigma.weight" in sp and "h_s.mlp_sigma.0.weight" not in sp


def test_default_init_matches_reference_gdn(golden):
    g = sic.GDN(5)
    assert torch.equal(g.beta, torch.sqrt(torch.ones(5) + 2 ** -18))
    assert torch.equal(g.gamma_conv.weight.view(-1), torch.sqrt(torch.full((5,), 0.1) + 2 ** -18))
    assert g.gamma.shape == (5, 5)


def test_quantize_static_and_errors():
    x = torch.tensor([0.5, 1.5, 2.5, -0.5, -0.4])
    with pytest.raises(sic.SicError):                       # no CPU path: the static helper runs kernel K1 like forward() does
        sic.CompressionModel.quantize(x, "round")
    with pytest.raises(ValueError):
        sic.CompressionModel.quantize(x, "floor")
    with pytest.raises(ValueError):
        sic.rate_distortion_loss({"nll_y": x, "nll_z": x, "x_hat": x.view(1, 1, 1, 5)}, x.view(1, 1, 1, 5), dist="psnr")


def test_loss_falls_back_to_tensor_sum_like_reference():
    out = {"nll_y": torch.full((2, 3, 4, 4), 0.5), "nll_z": torch.full((2, 2, 1, 1), 0.25), "x_hat": torch.zeros(2, 3, 16, 16)}
    x = torch.ones(2, 3, 16, 16)
    loss, R, D = sic.rate_distortion_loss(out, x, lambda_rd=10.0, dist="mse")
    assert abs(float(R) - (48.0 + 1.0) / (2 * 16 * 16)) < 1e-7 and float(D) == 1.0 and abs(float(loss) - (10.0 + float(R))) < 1e-6
    out["nll_y"] = -out["nll_y"]
    _, R, _ = sic.rate_distortion_loss(out, x, dist="mse")
    assert float(R) == 0.0                                                  # clamp(min=0), model.py:79


# ------------------------------------------------------------------------------------------------------- entropy coder
def _random_case(rng, C, hw, L):
    sig = np.exp(rng.normal(0, 1, C)).astype(np.float32)
    nu = (2 + rng.random(C) * 20).astype(np.float32)
    mn = np.array([-(L // 2)], np.int32)
    tab = clib.build_tables("studentt", sig, nu, np.zeros(C, np.int32), mn, mn + L - 1)
    sym = np.clip(np.rint(rng.standard_t(3, (C, hw)) * sig[:, None]).astype(np.int64) - mn[0], 0, L - 1).astype(np.int32)
    return tab, sym


@pytest.mark.parametrize("C,hw,L", [(1, 1, 2), (3, 16, 21), (7, 33, 40), (8, 64, 59), (2, 1000, 31)])
def test_rans_host_matches_oracle_and_round_trips(C, hw, L):
    rng = np.random.default_rng(C * 1000 + hw)
    tab, sym = _random_case(rng, C, hw, L)
    ours = F.rans_encode(sym, tab, L, hw)
    assert ours == clib.rans_encode(sym, tab, L, hw)                         # identical bytes
    assert np.array_equal(F.rans_decode(ours, sym.size, tab, L, hw), sym.ravel())
    assert np.array_equal(clib.rans_decode(ours, sym.size, tab, L, hw), sym.ravel())


def test_rans_zero_width_symbols_are_still_codable():
    """pmf_to_uint16_cdf leaves tail symbols with zero width (pmf < 1/65535); the coder's widening keeps them decodable."""
    tab = np.array([[0, 0, 0, 65535, 65535, 65535]], np.uint16)           # only symbol 2 has mass
    sym = np.array([0, 1, 2, 3, 4, 2, 2, 0], np.int32)
    data = F.rans_encode(sym, tab, 5, sym.size)
    assert np.array_equal(F.rans_decode(data, sym.size, tab, 5, sym.size), sym)


def test_rans_golden_bytes():
    """Known-answer vector for format SIC-RANS-1 (first written by the oracle coder; guards the format against drift)."""
    tab = np.array([[0, 8192, 40000, 60000, 65535]], np.uint16)
    sym = np.array([1, 1, 2, 0, 3, 1, 1, 2] * 60, np.int32)
    data = clib.rans_encode(sym, tab, 4, sym.size)
    assert F.rans_encode(sym, tab, 4, sym.size) == data
    assert len(data) == 184 and data[:8].hex() == GOLDEN_HEAD and data[-8:].hex() == GOLDEN_TAIL


GOLDEN_HEAD = "008900bf008900bf"
GOLDEN_TAIL = "a0fca0fca0fca0fc"


def test_rans_errors():
    tab = np.array([[0, 30000, 65535]], np.uint16)
    with pytest.raises(sic.SicError):
        F.rans_encode(np.array([0, 2], np.int32), tab, 2, 2)                # symbol outside [0,L)
    data = F.rans_encode(np.array([0, 1] * 2000, np.int32), tab, 2, 4000)
    assert len(data) > 300
    with pytest.raises(sic.SicError):
        F.rans_decode(data[:100], 4000, tab, 2, 4000)                        # shorter than the state header
    with pytest.raises(sic.SicError):
        F.rans_decode(data[:-2], 4000, tab, 2, 4000)                         # truncated payload
    assert F.rans_encode(np.zeros(0, np.int32), tab, 2, 1) == bytes([0, 0, 1, 0] * 32)   # empty stream = 32 initial states


def test_rans_damage_is_detected_not_decoded():
    """A stream that is damaged but long enough used to decode silently to garbage.  The decoder now requires every lane to end
    in the coder's initial state (2^16) and every word to be consumed: flipped bits, stray words and a wrong symbol count
    are all reported (SIC_E_CORRUPT / SIC_E_TRUNCATED), never returned as symbols."""
    rng = np.random.default_rng(11)
    tab = np.array([[0, 9000, 30000, 52000, 65535]], np.uint16)
    sym = rng.integers(0, 4, 3000).astype(np.int32)
    data = F.rans_encode(sym, tab, 4, sym.size)
    assert np.array_equal(F.rans_decode(data, sym.size, tab, 4, sym.size), sym)
    for pos in (3, 77, 128, 129, len(data) // 2, len(data) - 1):         # state header, first word, middle, last word
        hurt = bytearray(data); hurt[pos] ^= 0x10
        with pytest.raises(sic.SicError):
            F.rans_decode(bytes(hurt), sym.size, tab, 4, sym.size)
    with pytest.raises(sic.SicError):
        F.rans_decode(data + b"\x00\x00", sym.size, tab, 4, sym.size)     # stray word after the stream
    with pytest.raises(sic.SicError):
        F.rans_decode(data, sym.size - 40, tab, 4, sym.size - 40)         # fewer symbols than were coded: words are left over


def test_container_rejects_impossible_supports():
    """unpack() checks the CRC, which anyone can recompute; min > max or an empty latent must still be refused before the
    supports reach the table builder and the device decoder (L <= 0 as uint32 walked out of the table row)."""
    from domain_specific_image_compression_b200 import container
    good = {"strings": [[b"z" * 130, b"y" * 140]], "shape_y": [1, 4, 2, 2], "shape_z": [1, 2, 1, 1],
            "min_y": [-3], "max_y": [4], "min_z": [-2], "max_z": [2]}
    assert container.unpack(container.pack(good))["max_y"] == [4]
    for key, val in (("max_y", [-4]), ("max_z", [-3]), ("max_y", [5000]), ("shape_y", [1, 0, 2, 2])):
        bad = dict(good); bad[key] = val
        with pytest.raises(container.ContainerError):
            container.unpack(container.pack(bad))


# ------------------------------------------------------------------------------------------------------- MS-SSIM
def _msssim_numpy(x, y, weights):
    """Independent float64 evaluation of the published MS-SSIM algorithm with scipy.ndimage (valid windows)."""
    k = np.arange(11) - 5.0
    g = np.exp(-k ** 2 / (2 * 1.5 ** 2)); g /= g.sum()
    def blur(a):
        a = ndimage.correlate1d(a, g, axis=-1, mode="constant")
        a = ndimage.correlate1d(a, g, axis=-2, mode="constant")
        return a[..., 5:-5, 5:-5]
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    vals = []
    for lvl in range(len(weights)):
        if lvl > 0:
            x = 0.25 * (x[..., 0::2, 0::2] + x[..., 1::2, 0::2] + x[..., 0::2, 1::2] + x[..., 1::2, 1::2])
            y = 0.25 * (y[..., 0::2, 0::2] + y[..., 1::2, 0::2] + y[..., 0::2, 1::2] + y[..., 1::2, 1::2])
        mx, my = blur(x), blur(y)
        sxx, syy, sxy = blur(x * x) - mx * mx, blur(y * y) - my * my, blur(x * y) - mx * my
        cs = (2 * sxy + c2) / (sxx + syy + c2)
        ss = (2 * mx * my + c1) / (mx * mx + my * my + c1) * cs
        vals.append((ss if lvl == len(weights) - 1 else cs).mean(axis=(-1, -2)))
    v = np.maximum(np.stack(vals), 0.0)
    return np.prod(v ** np.asarray(weights)[:, None, None], axis=0).mean(1).mean(0)


def _msssim_numpy_odd(x, y, weights):
    """_msssim_numpy with piq's handling of odd sizes: replicate-pad one row/column on the top/left before the 2x2 mean."""
    k = np.arange(11) - 5.0
    g = np.exp(-k ** 2 / (2 * 1.5 ** 2)); g /= g.sum()
    def blur(a):
        a = ndimage.correlate1d(a, g, axis=-1, mode="constant")
        a = ndimage.correlate1d(a, g, axis=-2, mode="constant")
        return a[..., 5:-5, 5:-5]
    def pool(a):
        pad = max(a.shape[-2] % 2, a.shape[-1] % 2)
        if pad:
            a = np.pad(a, [(0, 0)] * (a.ndim - 2) + [(pad, 0), (pad, 0)], mode="edge")
        h2, w2 = a.shape[-2] // 2 * 2, a.shape[-1] // 2 * 2
        a = a[..., :h2, :w2]
        return 0.25 * (a[..., 0::2, 0::2] + a[..., 1::2, 0::2] + a[..., 0::2, 1::2] + a[..., 1::2, 1::2])
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    vals = []
    for lvl in range(len(weights)):
        if lvl > 0:
            x, y = pool(x), pool(y)
        mx, my = blur(x), blur(y)
        sxx, syy, sxy = blur(x * x) - mx * mx, blur(y * y) - my * my, blur(x * y) - mx * my
        cs = (2 * sxy + c2) / (sxx + syy + c2)
        ss = (2 * mx * my + c1) / (mx * mx + my * my + c1) * cs
        vals.append((ss if lvl == len(weights) - 1 else cs).mean(axis=(-1, -2)))
    v = np.maximum(np.stack(vals), 0.0)
    return np.prod(v ** np.asarray(weights)[:, None, None], axis=0).mean(1).mean(0)


def test_oracle_msssim_restatement_matches_independent_evaluation_and_product_statement():
    """oracle/torch_port.multi_scale_ssim (what the CPU baseline arm runs, and what the GPU tests check the fused kernels
    against) versus the independent scipy evaluation."""
    from oracle import torch_port as TP
    rng = np.random.default_rng(1)
    w = [0.3, 0.5, 0.2]
    x = rng.random((2, 3, 64, 64))
    y = np.clip(x + 0.1 * rng.standard_normal(x.shape), 0, 1)
    xt, yt = torch.from_numpy(x).float(), torch.from_numpy(y).float()
    got = float(TP.multi_scale_ssim(xt, yt, 1.0, torch.tensor(w)))
    assert abs(got - _msssim_numpy(x, y, w)) < 2e-5
    assert abs(float(TP.multi_scale_ssim(xt, xt, 1.0, torch.tensor(w))) - 1.0) < 1e-6
    a = rng.random((1, 3, 77, 101))                                           # odd sizes: replicate-padded pooling
    b = np.clip(a * 0.9 + 0.05, 0, 1)
    got = float(TP.multi_scale_ssim(torch.from_numpy(a).float(), torch.from_numpy(b).float(), 1.0, torch.tensor(w)))
    assert abs(got - _msssim_numpy_odd(a, b, w)) < 2e-5
    with pytest.raises(ValueError):
        TP.multi_scale_ssim(torch.rand(1, 3, 32, 32), torch.rand(1, 3, 32, 32), 1.0, torch.tensor(w))


def test_product_loss_has_no_cpu_path():
    """The product's distortion term goes through the CUDA kernels unconditionally: CPU tensors raise (north_star: no CPU
    fallback); size validation still comes first, as in piq."""
    w = torch.tensor([0.3, 0.5, 0.2])
    with pytest.raises(sic.SicError):
        multi_scale_ssim(torch.rand(1, 3, 64, 64), torch.rand(1, 3, 64, 64), 1.0, w)
    with pytest.raises(ValueError):
        multi_scale_ssim(torch.rand(1, 3, 32, 32), torch.rand(1, 3, 32, 32), 1.0, w)   # < 41 px for 3 scales


def test_rgb_channel_padding_is_the_identity_and_keeps_parameter_shapes():
    """layers._rgb_padded (opt-in PAD_RGB_CHANNELS): zero-padding the 3-band axis of the first conv / last transposed conv
    changes nothing in exact arithmetic and routes the gradients back to the reference-shaped parameters."""
    import torch.nn as nn
    from domain_specific_image_compression_b200 import layers as L
    old = L.PAD_RGB_CHANNELS
    try:
        for pad_to in (4, 8):
            L.PAD_RGB_CHANNELS = pad_to
            torch.manual_seed(pad_to)
            c = nn.Conv2d(3, 8, 3, 1, 1)
            x = torch.rand(2, 3, 16, 16)
            t, b = L._rgb_padded(c, x)
            assert b is c.bias and torch.allclose(t + b.view(1, -1, 1, 1), c(x), atol=1e-6)
            tc, _ = L._rgb_padded(c, x.contiguous(memory_format=torch.channels_last))
            assert torch.allclose(tc, t, atol=1e-6)
            d = nn.ConvTranspose2d(8, 3, 5, 2, 2, output_padding=1)
            z = torch.rand(2, 8, 8, 8)
            y, b2 = L._rgb_padded(d, z)
            assert b2 is None and y.shape == (2, 3, 16, 16) and torch.allclose(y, d(z), atol=1e-6)
            y.sum().backward()
            assert d.weight.grad.shape == d.weight.shape and d.bias.grad.shape == (3,)
            assert L._rgb_padded(nn.Conv2d(8, 8, 3), torch.rand(1, 8, 8, 8)) is None       # other layers are left alone
    finally:
        L.PAD_RGB_CHANNELS = old
