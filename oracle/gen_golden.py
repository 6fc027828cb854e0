#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REFERENCE itself (imported from /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/gen_golden.py
The script, not its output, is the source of truth; the fixtures are committed so that tests run anywhere.

What is imported unmodified:  distributions.py, layers.py, model.py (with an empty `piq` stub module so that
`import piq` at model.py:6 succeeds; `dist="mse"` never touches it), and `pmf_to_uint16_cdf` / `gaussian_cdf` /
`custom_compress` from eval_selfcontained_entropy.py (with stub `torchac` / `pytorch_msssim` modules).
`custom_compress` cannot run as written; the "repaired" fixture patches, from outside,
(1) torch.floor/ceil on python floats -> math.floor/ceil, (2) StudentT.cdf -> scipy.special.stdtr in float64 cast to
float32, (3) a capturing fake torchac.encode_float_cdf.  No reference file is modified or copied.
"""
from __future__ import annotations

import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/code/modelv2"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference not mounted at /root/reference; fixtures can only be regenerated in the build container")
    sys.path.insert(0, REF)
    sys.modules.setdefault("piq", types.ModuleType("piq"))
    captured = []
    fake_ac = types.ModuleType("torchac")
    fake_ac.encode_float_cdf = lambda cdf, sym, **kw: (captured.append((np.array(cdf), np.array(sym))), b"")[1]
    fake_ac.decode_float_cdf = lambda *a, **k: None
    sys.modules["torchac"] = fake_ac
    pm = types.ModuleType("pytorch_msssim")
    pm.ms_ssim = lambda *a, **k: None
    sys.modules["pytorch_msssim"] = pm
    import distributions, layers, model, eval_selfcontained_entropy as ese  # noqa: E401
    return distributions, layers, model, ese, captured


def spread_init(m, seed):
    """SURVEY 8(d) 'spread' variant: default init gives y == 0 and nu == 2; widen the last analysis convs,
    and perturb GDN / prior parameters so that every parameter matters."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        m.g_a.g_a[14].weight.mul_(40.0)
        m.h_a.h_a[6].weight.mul_(40.0)
        m.h_s.mlp_sigma[2].bias.add_(torch.randn(m.h_s.mlp_sigma[2].bias.shape, generator=g) * 0.7)
        m.h_s.mlp_nu[2].bias.add_(1.5 + torch.randn(m.h_s.mlp_nu[2].bias.shape, generator=g))
        m.z_prior.log_sigma.add_(torch.randn(m.z_prior.log_sigma.shape, generator=g) * 0.5 + 0.5)
        for mod in m.modules():
            if mod.__class__.__name__ == "GDN":
                mod.beta.mul_(0.75 + 0.5 * torch.rand(mod.beta.shape, generator=g))
                mod.gamma_conv.weight.mul_(0.5 + torch.rand(mod.gamma_conv.weight.shape, generator=g))


def gen_likelihood(distributions):
    torch.manual_seed(1)
    B, C, h, w = 4, 24, 8, 8
    st = distributions.StudentT()
    x = (torch.randn(B, C, h, w) * 3).requires_grad_(True)
    # broadcast parameters incl. values outside both clamp ranges and exactly on the bounds
    sig = torch.exp(torch.randn(B, C, 1, 1) * 1.5)
    nu = torch.exp(torch.randn(B, C, 1, 1) + 1.5)
    sig.view(-1)[:6] = torch.tensor([1e-4, 1e-3, 1e3, 2e3, 0.5, 7.0])
    nu.view(-1)[:6] = torch.tensor([1.5, 2.0, 100.0, 150.0, 2.0001, 99.0])
    sig.requires_grad_(True), nu.requires_grad_(True)
    g = torch.randn(B, C, h, w)
    nll = st.neg_log2_prob(x, sig.expand(B, C, h, w), nu.expand(B, C, h, w))
    (nll * g).sum().backward()
    out = dict(x=x.detach(), sigma_bc=sig.detach(), nu_bc=nu.detach(), g=g, nll_bc=nll.detach(), dx_bc=x.grad.clone(),
               dsigma_bc=sig.grad.clone(), dnu_bc=nu.grad.clone())
    # spatial parameters
    x.grad = None
    sig_s = torch.exp(torch.randn(B, C, h, w) * 1.5).requires_grad_(True)
    nu_s = torch.exp(torch.randn(B, C, h, w) + 1.5).requires_grad_(True)
    nll = st.neg_log2_prob(x, sig_s, nu_s)
    (nll * g).sum().backward()
    out.update(sigma_sp=sig_s.detach(), nu_sp=nu_s.detach(), nll_sp=nll.detach(), dx_sp=x.grad.clone(),
               dsigma_sp=sig_s.grad.clone(), dnu_sp=nu_s.grad.clone())
    # Gaussian z prior
    Cz = 16
    fg = distributions.FactorizedGaussian(Cz)
    with torch.no_grad():
        fg.log_sigma.copy_(torch.randn(Cz) * 1.2)
        fg.log_sigma[:3] = torch.tensor([-8.0, 8.0, math.log(1e3)])
    z = (torch.randn(B, Cz, 4, 4) * 2).requires_grad_(True)
    gz = torch.randn(B, Cz, 4, 4)
    nz = fg.neg_log2_prob(z)
    (nz * gz).sum().backward()
    out.update(z=z.detach(), log_sigma_z=fg.log_sigma.detach(), gz=gz, nll_z=nz.detach(), dz=z.grad.clone(),
               dlog_sigma_z=fg.log_sigma.grad.clone())
    np.savez_compressed(os.path.join(OUT, "likelihood.npz"), **{k: v.numpy() for k, v in out.items()})


def gen_gdn(layers):
    torch.manual_seed(2)
    B, C, H, W = 2, 16, 12, 10
    out = {}
    for inverse in (False, True):
        m = layers.GDN(C, inverse=inverse)
        with torch.no_grad():
            m.beta.mul_(0.75 + 0.5 * torch.rand(C))
            m.gamma_conv.weight.mul_(0.5 + torch.rand(C, 1, 1, 1))
        x = (torch.randn(B, C, H, W) * 2.5).requires_grad_(True)
        g = torch.randn(B, C, H, W)
        y = m(x)
        (y * g).sum().backward()
        tag = "igdn" if inverse else "gdn"
        assert m.gamma.grad is None      # the CxC parameter is dead
        out.update({f"{tag}_x": x.detach(), f"{tag}_g": g, f"{tag}_beta": m.beta.detach(),
                    f"{tag}_weight": m.gamma_conv.weight.detach(), f"{tag}_y": y.detach(), f"{tag}_dx": x.grad.clone(),
                    f"{tag}_dbeta": m.beta.grad.clone(), f"{tag}_dweight": m.gamma_conv.weight.grad.clone()})
    np.savez_compressed(os.path.join(OUT, "gdn.npz"), **{k: v.numpy() for k, v in out.items()})


def gen_model(model):
    torch.manual_seed(42)                                    # config.py:32
    N, M = 16, 24
    m = model.CompressionModel(N=N, M=M, spatial_params=False, min_nu=2.0, max_nu=100.0)
    spread_init(m, 7)
    x = torch.nn.functional.interpolate(torch.rand(2, 3, 16, 16), size=(64, 64), mode="bilinear").clamp(0, 1)
    out = {"x": x}
    for k, v in m.state_dict().items():
        out["sd." + k] = v
    m.eval()
    with torch.no_grad():
        o = m(x, quant_mode="round")
        loss, R, D = model.rate_distortion_loss(o, x, lambda_rd=100.0, dist="mse")
    for k, v in o.items():
        out["eval." + k] = v.contiguous()
    out.update({"eval.loss": loss, "eval.R": R, "eval.D": D})
    m.train()
    torch.manual_seed(123)
    o = m(x, quant_mode="noise")
    loss, R, D = model.rate_distortion_loss(o, x, lambda_rd=100.0, dist="mse")
    loss.backward()
    torch.manual_seed(123)                                   # replay the two uniform_ draws of model.py:44-45
    ny = torch.empty_like(o["y"]).uniform_(-0.5, 0.5)
    nz = torch.empty_like(o["z"]).uniform_(-0.5, 0.5)
    assert torch.equal(o["y"] + ny, o["y_tilde"]) and torch.equal(o["z"] + nz, o["z_tilde"])
    for k, v in o.items():
        out["train." + k] = v.detach().contiguous()
    out.update({"train.noise_y": ny, "train.noise_z": nz, "train.loss": loss.detach(), "train.R": R, "train.D": D})
    for k, p in m.named_parameters():
        if p.grad is not None:
            out["grad." + k] = p.grad
    np.savez_compressed(os.path.join(OUT, "model_small.npz"), **{k: v.detach().numpy() for k, v in out.items()})
    return m


def gen_tables(model, ese, captured, m):
    # (a) the reference's own pmf_to_uint16_cdf, unmodified
    torch.manual_seed(3)
    out = {}
    for i, shape in enumerate([(21, 8, 1, 1), (37, 5, 2, 3), (64, 3, 1, 1)]):
        pmf = torch.rand(*shape) ** 4 + 1e-12
        pmf = pmf / pmf.sum(dim=0, keepdim=True)
        out[f"pmf{i}"] = pmf
        out[f"cdf{i}"] = torch.from_numpy(ese.pmf_to_uint16_cdf(pmf).astype(np.int32))
    np.savez_compressed(os.path.join(OUT, "pmf_to_cdf.npz"), **{k: v.numpy() for k, v in out.items()})

    # (b) repaired custom_compress: tables + symbols as the reference intends them
    from scipy import special as sp

    class _TorchProxy:
        """Forwards to torch but lets floor/ceil accept the python floats the script passes (defect 1)."""
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def floor(v):
            return math.floor(v) if isinstance(v, float) else torch.floor(v)

        @staticmethod
        def ceil(v):
            return math.ceil(v) if isinstance(v, float) else torch.ceil(v)

    ese.torch = _TorchProxy()

    def scipy_cdf(self, value):                              # defect 2: torch has no StudentT.cdf
        t = ((value - self.loc) / self.scale).to(torch.float32)
        t, df = torch.broadcast_tensors(t, self.df)
        return torch.from_numpy(sp.stdtr(df.double().numpy(), t.double().numpy())).to(torch.float32)

    torch.distributions.StudentT.cdf = scipy_cdf
    m.eval()
    x = torch.nn.functional.interpolate(torch.rand(2, 3, 32, 32), size=(128, 128), mode="bilinear").clamp(0, 1)
    with torch.no_grad():
        o = m(x, quant_mode="round")
        captured.clear()
        comp = ese.custom_compress(m, x, tail=10)
    res = {"x": x, "y_q": o["y_tilde"], "z_q": o["z_tilde"], "sigma": o["sigma"][:, :, 0, 0].contiguous(),
           "nu": o["nu"][:, :, 0, 0].contiguous(), "log_sigma_z": m.z_prior.log_sigma.detach(),
           "min_y": torch.tensor(comp["min_y"]), "max_y": torch.tensor(comp["max_y"]),
           "min_z": torch.tensor(comp["min_z"]), "max_z": torch.tensor(comp["max_z"])}
    res = {k: v.numpy() for k, v in res.items()}
    for b in range(x.size(0)):
        cz, sz = captured[2 * b]
        cy, sy = captured[2 * b + 1]
        assert (cy == cy[:, :, :1, :1]).all()                # replicated over (h,w) in broadcast mode
        res[f"cdf_z{b}"] = cz[:, :, 0, 0].T.astype(np.int32).copy()      # -> [C, L+1], symbol axis last
        res[f"cdf_y{b}"] = cy[:, :, 0, 0].T.astype(np.int32).copy()
        res[f"sym_z{b}"] = sz.astype(np.int32)
        res[f"sym_y{b}"] = sy.astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "tables_repaired.npz"), **res)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    distributions, layers, model, ese, captured = import_reference()
    gen_likelihood(distributions)
    gen_gdn(layers)
    m = gen_model(model)
    gen_tables(model, ese, captured, m)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
