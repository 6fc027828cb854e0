"""CompressionModel / rate_distortion_loss with the API of /root/reference/code/modelv2/model.py, plus compress() /
decompress() methods that carry the semantics of custom_compress / custom_decompress
(/root/reference/code/modelv2/eval_selfcontained_entropy.py:26-123).

Drop-in contract: constructor signature, `forward(x, quant_mode)` returning exactly the nine keys
x_hat, nll_y, nll_z, y, y_tilde, z, z_tilde, sigma, nu; static `quantize`; `rate_distortion_loss(out, x, lambda_rd, dist)`
returning (loss, R.detach(), D.detach()); 90 state_dict keys identical to the reference's.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as F_sic
from .container import ContainerError, validate_supports
from .distributions import FactorizedGaussian, StudentT
from .layers import AnalysisTransform, HyperAnalysis, HyperSynthesis, SynthesisTransform
from .losses import multi_scale_ssim


# Opt-in (bench.py switches it on): in TRAINING mode run the hyperprior branch (h_a -> K1(z) -> h_s -> tail -> K1 likelihood of y)
# on a side stream next to the synthesis transform.  The two branches only share y / y_tilde: g_s(y_tilde) does not need sigma / nu,
# and the likelihood does not need x_hat.  The hyper branch is ~40 launches on 16x16 ... 4x4 maps (latency-bound, a few CTAs each)
# which otherwise sit in line between the large analysis and synthesis kernels, forward and backward (autograd runs each backward
# node on its forward's stream).  Same values as the sequential schedule: y_tilde is produced first by the quantise-only form of K1
# (same arithmetic), the likelihood kernel then takes y_tilde with quant = none.
OVERLAP_HYPER_BRANCH = False
_side_streams = {}


def side_stream(device) -> "torch.cuda.Stream":
    """The per-device side stream of the overlapped schedule (created on first use)."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _side_streams.get(idx)
    if st is None:
        st = _side_streams[idx] = torch.cuda.Stream(device=idx)
    return st


def join_side_streams(device) -> None:
    """Make the current stream wait for everything queued on the side stream so far (no-op if the overlapped schedule never ran).
    FlatTrainer calls it before it packs a gradient bucket: the bucket's gradients may have been produced on either stream."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _side_streams.get(idx)
    if st is not None:
        torch.cuda.current_stream(idx).wait_stream(st)


FUSE_HYPER_TAIL = True      # False: run the hyper-synthesis tail as the reference's eager op chain (layers.py:141-152, model.py:54-55)


class CompressionModel(nn.Module):
    def __init__(self, N=128, M=192, spatial_params=False, min_nu=1.1, max_nu=100.0, likelihood: str = "density"):
        super().__init__()
        self.g_a = AnalysisTransform(N, M)
        self.g_s = SynthesisTransform(N, M)
        self.h_a = HyperAnalysis(M, N)
        self.h_s = HyperSynthesis(N, M, spatial_params=spatial_params)
        self.studentT = StudentT(mode=likelihood)
        self.z_prior = FactorizedGaussian(N)
        self.min_nu = min_nu
        self.max_nu = max_nu
        self.spatial_params = spatial_params
        self.likelihood = likelihood          # 'density' (what the reference trains with) | 'cdf_diff' (north_star variant)

    @staticmethod
    def quantize(x, mode):
        """model.py:27-35.  Standalone helper kept for API parity (forward() fuses quantisation into the same kernel together with
        the likelihood).  Runs kernel K1 in quantise-only form — CUDA float32 only, like every op here; 'noise' draws from the
        in-kernel Philox stream seeded by torch.manual_seed, not from torch's generator."""
        if mode not in ("noise", "round"):
            raise ValueError(f"Unknown quant mode: {mode}")
        return F_sic.quantize(x, mode)

    def _student_params(self, z_tilde, like):
        """model.py:47-55: hyper-synthesis + sigma/nu post-processing.  Returns kernel-layout and dict-layout tensors."""
        if not self.spatial_params and FUSE_HYPER_TAIL:
            # N4: pool -> two MLPs -> exp / clamp as ONE kernel with a fixed, batch-size-independent summation order (the eager
            # chain is ~15 launches, and its cuDNN/cuBLAS GEMMs may round differently at different batch sizes, which would
            # desynchronise encoder and decoder tables)
            t = self.h_s.trunk(z_tilde)
            sigma_k, nu_k = F_sic.hyper_tail(t, self.h_s.mlp_sigma, self.h_s.mlp_nu, self.min_nu, self.max_nu)
            return sigma_k, nu_k, sigma_k.expand_as(like), nu_k.expand_as(like)
        log_sigma, log_nu = self.h_s(z_tilde)
        if self.spatial_params:
            sigma = torch.exp(log_sigma)
            nu = torch.clamp(torch.exp(log_nu), min=self.min_nu, max=self.max_nu)
            return sigma, nu, sigma, nu
        sigma_k = torch.exp(log_sigma).mean(dim=(2, 3), keepdim=True)                                   # [B,M,1,1]
        nu_k = torch.clamp(torch.exp(log_nu).mean(dim=(2, 3), keepdim=True), self.min_nu, self.max_nu)
        return sigma_k, nu_k, sigma_k.expand_as(like), nu_k.expand_as(like)

    def forward(self, x, quant_mode="noise", noise_y=None, noise_z=None, synthesize=True):
        """noise_y / noise_z (optional, not in the reference): supply the uniform draws for bit-exact parity runs.
        synthesize=False (not in the reference; compress() uses it): stop after the entropy bottleneck - the same dict without
        `x_hat`; the encoder needs latents, sigma and nu, not the reconstruction (g_s is ~40 % of a forward pass)."""
        if quant_mode not in ("noise", "round"):
            raise ValueError(f"Unknown quant mode: {quant_mode}")
        y = self.g_a(x)
        if OVERLAP_HYPER_BRANCH and self.training and quant_mode == "noise" and y.is_cuda and synthesize:
            return self._forward_overlapped(y, noise_y, noise_z)
        z = self.h_a(y)
        # K1 (Gaussian): quantise z + nll_z + per-patch bits in one launch (model.py:45,59)
        z_tilde, nll_z, bits_z = F_sic.bottleneck(z, self.z_prior.log_sigma, quant=quant_mode, lik="gaussian", noise=noise_z)
        sigma_k, nu_k, sigma, nu = self._student_params(z_tilde, y)
        # K1 (Student-t): quantise y + nll_y + per-patch bits in one launch (model.py:44,58)
        y_tilde, nll_y, bits_y = F_sic.bottleneck(y, sigma_k, nu_k, quant=quant_mode, lik=self.likelihood, noise=noise_y)
        nll_y._sic_bits, nll_z._sic_bits = bits_y, bits_z
        if self.training:
            y_hat = y_tilde                                                 # model.py:62
        else:
            y_hat = y_tilde if quant_mode == "round" else torch.round(y)   # round(y) twice in the reference; identical bits
        out = {"nll_y": nll_y, "nll_z": nll_z, "y": y, "y_tilde": y_tilde, "z": z, "z_tilde": z_tilde, "sigma": sigma, "nu": nu}
        if synthesize:
            out = {"x_hat": self.g_s(y_hat), **out}
        return out

    def _forward_overlapped(self, y, noise_y, noise_z):
        """Training forward with the hyperprior branch on a side stream (OVERLAP_HYPER_BRANCH).  Same nine outputs."""
        cur = torch.cuda.current_stream(y.device)
        side = side_stream(y.device)
        # y_tilde first (model.py:44), so that the synthesis transform can start: quantise-only K1, or the supplied draw
        y_tilde = (y + noise_y) if noise_y is not None else F_sic.quantize(y, "noise")
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            z = self.h_a(y)
            z_tilde, nll_z, bits_z = F_sic.bottleneck(z, self.z_prior.log_sigma, quant="noise", lik="gaussian", noise=noise_z)
            sigma_k, nu_k, sigma, nu = self._student_params(z_tilde, y)
            _, nll_y, bits_y = F_sic.bottleneck(y_tilde, sigma_k, nu_k, quant="none", lik=self.likelihood)
            nll_y._sic_bits, nll_z._sic_bits = bits_y, bits_z
        for t in (y, y_tilde):
            t.record_stream(side)
        x_hat = self.g_s(y_tilde)
        cur.wait_stream(side)
        for t in (z, z_tilde, nll_z, nll_y, bits_y, bits_z, sigma_k, nu_k):
            t.record_stream(cur)
        return {"x_hat": x_hat, "nll_y": nll_y, "nll_z": nll_z, "y": y, "y_tilde": y_tilde, "z": z, "z_tilde": z_tilde,
                "sigma": sigma, "nu": nu}

    # ------------------------------------------------------------------------------------------------ entropy coding
    @torch.no_grad()
    def compress(self, x, tail=10, coder="gpu"):
        """custom_compress (eval_selfcontained_entropy.py:26-74): same return dict.
        One host sync for the whole batch (the reference has four per patch); compact [B,C,L+1] tables instead of tables
        replicated over (h,w); coder='gpu' codes every stream of the batch concurrently on the device (one warp per stream),
        coder='host' uses the C++ coder — both emit identical bytes (format SIC-RANS-1)."""
        if coder not in ("gpu", "host"):
            raise ValueError("coder must be 'gpu' or 'host'")
        was_training = self.training
        self.eval()
        try:
            with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
                out = self(x, quant_mode="round", synthesize=False)    # the encoder does not need the reconstruction
        finally:
            self.train(was_training)
        y_q, z_q = out["y_tilde"], out["z_tilde"]
        B = x.size(0)
        sym_z, min_z, max_z = F_sic.quantize_indices(z_q, do_round=False, tail=tail)
        sym_y, min_y, max_y = F_sic.quantize_indices(y_q, do_round=False, tail=tail)
        ext = torch.stack([min_z, max_z, min_y, max_y]).cpu().numpy()       # the single device->host sync
        mnz, mxz, mny, mxy = (ext[i] for i in range(4))
        tab_z = F_sic.build_cdf_tables("gaussian", self.z_prior.log_sigma.detach(), None, B, min_z, max_z, int((mxz - mnz).max()) + 2)
        sig, nu = self._table_params(out["sigma"], out["nu"])
        tab_y = F_sic.build_cdf_tables("studentt", sig, nu, B, min_y, max_y, int((mxy - mny).max()) + 2, channels=y_q.size(1))
        spr_z = z_q.size(2) * z_q.size(3)
        spr_y = 1 if self.spatial_params else y_q.size(2) * y_q.size(3)
        if coder == "gpu":
            bz, nz = F_sic.rans_encode_device(sym_z.view(B, -1), tab_z, max_z - min_z + 1, spr_z, tab_z.shape[0] // B)
            by, ny = F_sic.rans_encode_device(sym_y.view(B, -1), tab_y, max_y - min_y + 1, spr_y, tab_y.shape[0] // B)
            # the output buffers are sized for the worst case (128 + 2 bytes per symbol: 6.3 MB for a 16 x 512^2 batch); copy the lengths
            # first (one tiny sync), then only the bytes that were written (~0.7 MB) - the full-capacity pageable copies cost ~3 ms
            nzy = torch.stack([nz, ny]).cpu().numpy()
            nz, ny = nzy[0], nzy[1]
            if (nz < 0).any() or (ny < 0).any():
                raise F_sic._lib.SicError("rANS encoder: symbol outside its table support")
            bz = bz[:, :int(nz.max())].contiguous().cpu().numpy()
            by = by[:, :int(ny.max())].contiguous().cpu().numpy()
            strings = [[bz[b, :nz[b]].tobytes(), by[b, :ny[b]].tobytes()] for b in range(B)]
        else:
            sym_z_h, sym_y_h = sym_z.cpu().numpy(), sym_y.cpu().numpy()
            tab_z_h = tab_z.cpu().numpy().reshape(B, -1, tab_z.shape[-1])
            tab_y_h = tab_y.cpu().numpy().reshape(B, -1, tab_y.shape[-1])
            strings = [[F_sic.rans_encode(sym_z_h[b], tab_z_h[b], int(mxz[b] - mnz[b] + 1), spr_z),
                        F_sic.rans_encode(sym_y_h[b], tab_y_h[b], int(mxy[b] - mny[b] + 1), spr_y)] for b in range(B)]
        return {"strings": strings, "shape_y": list(y_q.shape), "shape_z": list(z_q.shape),
                "min_y": [int(v) for v in mny], "max_y": [int(v) for v in mxy],
                "min_z": [int(v) for v in mnz], "max_z": [int(v) for v in mxz]}

    def _table_params(self, sigma, nu):
        if self.spatial_params:
            return sigma.contiguous().view(-1), nu.contiguous().view(-1)
        return sigma[:, :, 0, 0].contiguous().view(-1), nu[:, :, 0, 0].contiguous().view(-1)

    @staticmethod
    def _pack_streams(strings, which, n_sym, device):
        """byte strings of one latent -> uint8 [B, cap] (zero padded, cap % 4 == 0) + int32 lengths, on the device."""
        B = len(strings)
        # row stride = the longest stream (not the encoder's worst case of 128 + 2 bytes per symbol: that made this a 6 MB zero-fill
        # and pageable host->device copy for ~0.7 MB of payload)
        cap = (max(128, max(len(s[which]) for s in strings)) + 3) // 4 * 4
        buf = np.zeros((B, cap), np.uint8)
        lens = np.zeros(B, np.int32)
        for b in range(B):
            raw = np.frombuffer(strings[b][which], np.uint8)
            buf[b, :raw.size] = raw
            lens[b] = raw.size
        return torch.from_numpy(buf).to(device), torch.from_numpy(lens).to(device)

    def _decode(self, strings, which, n_sym, tables, mins, maxs, spr, coder, dev):
        """-> float32 latent [B, n_sym] on the device (symbol + min, eval_selfcontained_entropy.py:97,117)."""
        B = len(strings)
        if coder == "gpu":
            data, lens = self._pack_streams(strings, which, n_sym, dev)
            mins_d, maxs_d = torch.from_numpy(mins).to(dev), torch.from_numpy(maxs).to(dev)
            sym, status = F_sic.rans_decode_device(data, lens, tables, maxs_d - mins_d + 1, n_sym, spr, tables.shape[0] // B)
            st = status.cpu().numpy()                                         # one sync; erasures must not pass silently
            if st.any():
                bad = int(np.flatnonzero(st)[0])
                why = {-5: "truncated stream", -6: "damaged stream (final coder state / word count do not close)",
                       -1: "support outside what the tables hold"}.get(int(st[bad]), f"status {int(st[bad])}")
                raise F_sic._lib.SicError(f"rANS decoder: patch {bad}, stream {which}: {why}")
            return (sym + mins_d.view(B, 1)).to(torch.float32)
        tab_h = tables.cpu().numpy().reshape(B, -1, tables.shape[-1])
        lat = np.empty((B, n_sym), np.float32)
        for b in range(B):
            s = F_sic.rans_decode(strings[b][which], n_sym, tab_h[b], int(maxs[b] - mins[b] + 1), spr)
            lat[b] = (s + mins[b]).astype(np.float32)
        return torch.from_numpy(lat).to(dev)

    @torch.no_grad()
    def decompress(self, compressed, coder="gpu"):
        """custom_decompress (eval_selfcontained_entropy.py:76-123): returns x_hat.clamp(0,1) [B,3,H,W]."""
        if coder not in ("gpu", "host"):
            raise ValueError("coder must be 'gpu' or 'host'")
        dev = next(self.parameters()).device
        strings = compressed["strings"]
        shape_y, shape_z = list(compressed["shape_y"]), list(compressed["shape_z"])
        B = len(strings)
        if B == 0:
            raise F_sic._lib.SicError("decompress: no patches")
        try:
            validate_supports(compressed, B)                                  # min <= max, support <= 4096, non-empty shapes
        except ContainerError as e:
            raise F_sic._lib.SicError(f"decompress: {e}") from None
        mnz, mxz = np.asarray(compressed["min_z"], np.int32), np.asarray(compressed["max_z"], np.int32)
        mny, mxy = np.asarray(compressed["min_y"], np.int32), np.asarray(compressed["max_y"], np.int32)
        to_dev = lambda a: torch.from_numpy(a).to(dev)
        tab_z = F_sic.build_cdf_tables("gaussian", self.z_prior.log_sigma.detach(), None, B, to_dev(mnz), to_dev(mxz),
                                       int((mxz - mnz).max()) + 2)                                      # :88-94
        n_z = shape_z[1] * shape_z[2] * shape_z[3]
        z_hat = self._decode(strings, 0, n_z, tab_z, mnz, mxz, shape_z[2] * shape_z[3], coder, dev).view(B, *shape_z[1:])
        like = torch.empty(B, *shape_y[1:], device=dev)
        was_training = self.training
        self.eval()
        try:
            with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
                _, _, sigma, nu = self._student_params(z_hat, like)           # :99-106
                sig, nu = self._table_params(sigma, nu)
                tab_y = F_sic.build_cdf_tables("studentt", sig, nu, B, to_dev(mny), to_dev(mxy), int((mxy - mny).max()) + 2,
                                               channels=shape_y[1])           # :108-114
                n_y = shape_y[1] * shape_y[2] * shape_y[3]
                spr_y = 1 if self.spatial_params else shape_y[2] * shape_y[3]
                y_hat = self._decode(strings, 1, n_y, tab_y, mny, mxy, spr_y, coder, dev).view(B, *shape_y[1:])
                x_hat = self.g_s(y_hat)                                       # :119-120
        finally:
            self.train(was_training)
        return x_hat.clamp(0, 1)                                             # :123


_MSSSIM_WEIGHTS = {}


def _msssim_weights(device) -> torch.Tensor:
    """[0.3, 0.5, 0.2] of model.py:100, created once per device (a host->device copy per call would break graph capture)."""
    w = _MSSSIM_WEIGHTS.get(device)
    if w is None:
        w = torch.tensor([0.3, 0.5, 0.2], device=device)
        _MSSSIM_WEIGHTS[device] = w
    return w


def _patch_bits(t: torch.Tensor) -> torch.Tensor:
    """Bit counts whose sum is the sum of a nll map (model.py:77): the per-patch counts that kernel K1 already reduced
    (deterministically) when the map came out of this package's forward(), else the map's own sum as a one-element tensor."""
    bits = getattr(t, "_sic_bits", None)
    return bits if bits is not None else t.sum().reshape(1)


def rate_distortion_loss(out: Dict[str, torch.Tensor], x, lambda_rd=10000.0, dist="mssim"):
    """model.py:75-107."""
    N, C, H, W = x.shape
    if dist == "mse":
        d_val, similarity = F.mse_loss(out["x_hat"], x), False
    elif dist == "msssim":
        x_hat = out["x_hat"]
        if x_hat.shape[2:] != x.shape[2:]:
            x_hat = F.interpolate(x_hat, size=x.shape[2:], mode="bilinear", align_corners=False)
        # the clamp(0, 1) of model.py:98 happens inside the kernels
        d_val, similarity = multi_scale_ssim(x_hat, x, data_range=1.0, scale_weights=_msssim_weights(x.device), clamp01=True), True
    else:
        raise ValueError("dist must be 'mse' or 'msssim'")
    by, bz = _patch_bits(out["nll_y"]), _patch_bits(out["nll_z"])
    if d_val.is_cuda and d_val.dtype == torch.float32 and by.dtype == torch.float32 and bz.dtype == torch.float32:
        return F_sic.rd_loss_tail(by, bz, d_val, N * H * W, lambda_rd, similarity)       # one launch (and one in the backward)
    R = torch.clamp((by.sum() + bz.sum()) / (N * H * W), min=0.0)      # host tensors (unit tests of the formula): the reference's op chain
    D = 1.0 - d_val if similarity else d_val
    loss = lambda_rd * D + R
    return loss, R.detach(), D.detach()
