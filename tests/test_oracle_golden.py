"""The oracle against the REFERENCE's own outputs (tests/golden/*.npz, produced by oracle/gen_golden.py which imports
/root/reference).  This is what pins the oracle; the GPU tests then compare the CUDA path with the oracle."""
import numpy as np

from oracle import clib
from oracle import numpy_ref as R


def test_studentt_density_broadcast(golden):
    G = golden("likelihood")
    nll = R.studentt_nll_f32(G["x"], G["sigma_bc"], G["nu_bc"])
    ref = G["nll_bc"]
    # same fp32 op order; only lgamma/log/log1p implementations differ (scipy vs torch)
    np.testing.assert_allclose(nll, ref, rtol=1e-5, atol=1e-4)
    truth = R.studentt_nll_f64(G["x"], G["sigma_bc"], G["nu_bc"])
    np.testing.assert_allclose(truth, ref, rtol=1e-5, atol=1e-4)
    assert abs(truth.sum() - ref.astype(np.float64).sum()) <= 1e-6 * abs(truth.sum())


def test_studentt_density_spatial(golden):
    G = golden("likelihood")
    truth = R.studentt_nll_f64(G["x"], G["sigma_sp"], G["nu_sp"])
    np.testing.assert_allclose(truth, G["nll_sp"], rtol=1e-5, atol=1e-4)


def test_studentt_gradients(golden):
    G = golden("likelihood")
    dx, ds, dn = R.studentt_nll_grads_f64(G["x"], G["sigma_bc"], G["nu_bc"], G["g"])
    np.testing.assert_allclose(dx, G["dx_bc"], rtol=1e-4, atol=1e-5 * np.abs(G["dx_bc"]).max())
    np.testing.assert_allclose(ds.sum((2, 3), keepdims=True), G["dsigma_bc"], rtol=1e-4, atol=1e-5 * np.abs(G["dsigma_bc"]).max())
    # the reference's fp32 autograd of lgamma is noisy (1% at nu=100): loose absolute tolerance, scale = gradient magnitude
    np.testing.assert_allclose(dn.sum((2, 3), keepdims=True), G["dnu_bc"], rtol=2e-2, atol=2e-4)
    # clamp masks: entries 0 and 3 of the fixture sit outside the clamp ranges -> exactly zero (closed-interval rule)
    assert G["dsigma_bc"].ravel()[0] == 0 and ds.sum((2, 3)).ravel()[0] == 0
    assert G["dnu_bc"].ravel()[3] == 0 and dn.sum((2, 3)).ravel()[3] == 0
    assert G["dsigma_bc"].ravel()[1] != 0 and ds.sum((2, 3)).ravel()[1] != 0     # exactly on the bound: gradient passes
    dx, ds, dn = R.studentt_nll_grads_f64(G["x"], G["sigma_sp"], G["nu_sp"], G["g"])
    np.testing.assert_allclose(dx, G["dx_sp"], rtol=1e-4, atol=1e-5 * np.abs(G["dx_sp"]).max())
    np.testing.assert_allclose(ds, G["dsigma_sp"], rtol=1e-3, atol=1e-5 * np.abs(G["dsigma_sp"]).max())
    np.testing.assert_allclose(dn, G["dnu_sp"], rtol=2e-2, atol=2e-4)


def test_gaussian(golden):
    G = golden("likelihood")
    nll = R.gaussian_nll_f32(G["z"], G["log_sigma_z"])
    np.testing.assert_allclose(nll, G["nll_z"], rtol=1e-5, atol=1e-4)
    dz, dls = R.gaussian_nll_grads_f64(G["z"], G["log_sigma_z"], G["gz"])
    np.testing.assert_allclose(dz, G["dz"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dls, G["dlog_sigma_z"], rtol=1e-4, atol=1e-3)
    assert G["dlog_sigma_z"][0] == 0 and G["dlog_sigma_z"][1] == 0     # sigma outside [1e-3,1e3]: clamp kills the gradient


def test_gdn_forward_within_torch_cpu_sqrt_error(golden):
    """numpy (IEEE) replay vs the reference on CPU.  torch-CPU's sqrt goes through MKL VML (HA mode, <1 ulp but not correctly
    rounded) so ~0.5% of elements differ by 1-2 ulp; the bit-exact claim is against torch CUDA eager (tests/test_gpu_gdn.py)."""
    D = golden("gdn")
    for tag, inv in (("gdn", False), ("igdn", True)):
        y = R.gdn_diag_fwd_f32(D[tag + "_x"], D[tag + "_beta"], D[tag + "_weight"], inv)
        ulp = np.abs(y.view(np.int32) - D[tag + "_y"].view(np.int32))
        assert ulp.max() <= 2
        assert (ulp == 0).mean() > 0.98


def test_gdn_backward(golden):
    D = golden("gdn")
    for tag, inv in (("gdn", False), ("igdn", True)):
        dx, db, dw = R.gdn_diag_bwd_f64(D[tag + "_x"], D[tag + "_g"], D[tag + "_beta"], D[tag + "_weight"], inv)
        np.testing.assert_allclose(dx, D[tag + "_dx"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(db, D[tag + "_dbeta"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(dw, D[tag + "_dweight"].ravel(), rtol=1e-4, atol=1e-4)


def test_gdn_dense_equals_diag_for_diagonal_gamma(golden):
    D = golden("gdn")
    beta, gamma = R.gdn_effective_params(D["gdn_beta"], D["gdn_weight"])
    y_dense = R.gdn_dense_fwd_f64(D["gdn_x"], beta, np.diag(gamma.astype(np.float64)))
    np.testing.assert_allclose(y_dense, D["gdn_y"], rtol=1e-5, atol=1e-6)
    dx, db, dg = R.gdn_dense_bwd_f64(D["gdn_x"], D["gdn_g"], beta, np.diag(gamma.astype(np.float64)))
    np.testing.assert_allclose(dx, D["gdn_dx"], rtol=1e-4, atol=1e-5)


def test_pmf_to_uint16_cdf_bit_exact(golden):
    """T1: the reference's own pmf_to_uint16_cdf (eval_selfcontained_entropy.py:17-23), imported unmodified."""
    P = golden("pmf_to_cdf")
    for i in range(3):
        c = R.pmf_to_uint16_cdf_spec(P[f"pmf{i}"])
        assert np.array_equal(c.astype(np.int32), P[f"cdf{i}"])


def test_tables_vs_repaired_reference(golden):
    """T2/T3/I1 against the 'repaired' run of custom_compress (PARITY UNPINNED by the reference itself, see gen_golden.py).
    Symbols and supports are exact; table entries may differ by 1 LSB where torch's fp32 sum / erf rounds differently."""
    T = golden("tables_repaired")
    B, C = T["sigma"].shape
    total = mism = 0
    for b in range(B):
        sym, mn, mx = R.symbols_and_support(T["y_q"][b:b + 1], tail=10)
        assert np.array_equal(sym[0], T[f"sym_y{b}"]) and mn[0] == T["min_y"][b] and mx[0] == T["max_y"][b]
        sym, mn, mx = R.symbols_and_support(T["z_q"][b:b + 1], tail=10)
        assert np.array_equal(sym[0], T[f"sym_z{b}"]) and mn[0] == T["min_z"][b] and mx[0] == T["max_z"][b]
        ty = clib.build_tables("studentt", T["sigma"][b], T["nu"][b], np.zeros(C, np.int32), T["min_y"][b:b + 1], T["max_y"][b:b + 1])
        sz = clib.exp_f32(T["log_sigma_z"])
        tz = clib.build_tables("gaussian", sz, None, np.zeros(sz.size, np.int32), T["min_z"][b:b + 1], T["max_z"][b:b + 1])
        for mine, ref in ((ty, T[f"cdf_y{b}"]), (tz, T[f"cdf_z{b}"])):
            d = np.abs(mine.astype(np.int32) - ref)
            assert d.max() <= 1
            total += d.size
            mism += int((d != 0).sum())
            assert (mine[:, 0] == 0).all() and (mine[:, -1] == 65535).all()
            assert (np.diff(mine.astype(np.int32), axis=1) >= 0).all()
    assert mism <= 0.005 * total


def test_torch_port_matches_reference_model(golden):
    """The functional PyTorch port (CPU baseline / eager comparator) reproduces the imported reference bit for bit on CPU."""
    import torch
    from oracle import torch_port as TP
    torch.set_num_threads(1)
    G = golden("model_small")
    sd = {k[3:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("sd.")}
    x = torch.from_numpy(G["x"])
    with torch.no_grad():
        o = TP.forward(sd, x, "round", training=False)
    for k in ("y", "z", "y_tilde", "z_tilde", "sigma", "nu", "nll_y", "nll_z", "x_hat"):
        assert torch.equal(o[k], torch.from_numpy(G["eval." + k])), k
    with torch.no_grad():
        o = TP.forward(sd, x, "noise", training=True, noise_y=torch.from_numpy(G["train.noise_y"]), noise_z=torch.from_numpy(G["train.noise_z"]))
        loss, Rr, D = TP.loss_fn(o, x, 100.0, "mse")
    for k in ("y_tilde", "z_tilde", "nll_y", "nll_z", "x_hat"):
        assert torch.equal(o[k], torch.from_numpy(G["train." + k])), k
    assert float(loss) == float(G["train.loss"]) and float(Rr) == float(G["train.R"])
