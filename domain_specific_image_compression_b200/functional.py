"""torch.autograd front ends of the C-ABI kernels (include/sic.h).

PyTorch is plumbing here: it owns device memory and the stream; every number is produced by libsic.so.  Inputs must
be CUDA float32 tensors — there is no CPU or eager fallback (the CPU oracle lives in /oracle and is test-only).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (LIK_GAUSSIAN, LIK_STUDENTT_CDFDIFF, LIK_STUDENTT_DENSITY, PARAM_BROADCAST, PARAM_CHANNEL,
                   PARAM_SPATIAL, QUANT_NOISE_PHILOX, QUANT_NOISE_TENSOR, QUANT_NONE, QUANT_ROUND)

_QUANT = {"none": QUANT_NONE, "round": QUANT_ROUND, "noise": QUANT_NOISE_PHILOX}
_LIK = {"density": LIK_STUDENTT_DENSITY, "gaussian": LIK_GAUSSIAN, "cdf_diff": LIK_STUDENTT_CDFDIFF}

# ----------------------------------------------------------------------------------------------------------------------
# plumbing
_workspaces = {}
_philox = {}
launch_count = 0          # kernels launched through this module (bench.py reports it as gpu_launches)

# The reference trains under torch.cuda.amp.autocast + GradScaler (train.py:196-204, config.py TRAIN.amp=True): the cuDNN convs
# then hand float16 activations to GDN and to the likelihood.  The kernels compute in float32 (>= the reference's precision
# there), so every autograd front end casts its floating inputs up and runs with autocast off; the backward runs in the same
# state.  Without this the first GDN raised "expected float32" under the reference's default config.
_amp_fwd = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.SicError(f"{name}: expected a CUDA tensor — this package has no CPU path (the CPU oracle is oracle/, test only)")
    if t.dtype != torch.float32:
        raise _lib.SicError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def _dense_layout(t: torch.Tensor, name: str):
    """Returns (tensor, channels_last flag) without copying when `t` is dense in either NCHW or NHWC order."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.SicError(f"{name}: expected a CUDA tensor — this package has no CPU path (the CPU oracle is oracle/, test only)")
    if t.dtype != torch.float32:
        raise _lib.SicError(f"{name}: expected float32, got {t.dtype}")
    if t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last):
        return t, 1
    return t.contiguous(), 0


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _workspace(device: torch.device, nbytes: int, kind: str = "ticketed") -> torch.Tensor:
    """Zero-initialised scratch, one per (device, stream, kind).  'ticketed' buffers start with the retire counter of the
    K1 kernels, which those kernels leave at zero; 'scratch' buffers (GDN backward partials) are overwritten freely and
    must therefore never be the same allocation."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, kind)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def philox_state(device: torch.device) -> torch.Tensor:
    """Device-resident {seed, offset} for the in-kernel noise; re-seeded whenever torch.manual_seed changes."""
    seed = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        # data-parallel ranks share torch's seed; give every rank its own noise stream (rank 0 keeps the plain seed)
        seed = (seed + 0x9E3779B97F4A7C15 * torch.distributed.get_rank()) & 0x7FFFFFFFFFFFFFFF
    st = _philox.get(device.index)
    if st is None or st[0] != seed:
        t = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        _philox[device.index] = (seed, t)
        return t
    return st[1]


def _launch(rc: int, what: str) -> None:
    global launch_count
    _lib.check(rc, what)
    launch_count += 1


# ----------------------------------------------------------------------------------------------------------------------
# K1
class _Bottleneck(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, y, sigma, nu, mu, noise, quant_mode: int, lik_mode: int, layout: int, want_nll: bool):
        lib = _lib.load()
        y = _require_cuda_f32(y, "y")
        sigma = _require_cuda_f32(sigma, "sigma")
        nu = _require_cuda_f32(nu, "nu") if nu is not None else None
        mu = _require_cuda_f32(mu, "mu") if mu is not None else None
        B, C = y.shape[0], y.shape[1]
        HW = y.numel() // (B * C)
        n_par = {PARAM_BROADCAST: B * C, PARAM_SPATIAL: y.numel(), PARAM_CHANNEL: C}[layout]
        for t, nm in ((sigma, "sigma"), (nu, "nu"), (mu, "mu")):
            if t is not None and t.numel() != n_par:
                raise _lib.SicError(f"{nm}: {t.numel()} values, layout needs {n_par}")
        philox = None
        if quant_mode == QUANT_NOISE_TENSOR:
            noise = _require_cuda_f32(noise, "noise")
            if noise.shape != y.shape:
                raise _lib.SicError("noise must have the shape of y")
        elif quant_mode == QUANT_NOISE_PHILOX:
            philox = philox_state(y.device)
        y_tilde = torch.empty_like(y) if quant_mode != QUANT_NONE else None
        nll = torch.empty_like(y) if want_nll else None
        bits = torch.empty(B, dtype=torch.float32, device=y.device)
        nws = lib.sic_bottleneck_workspace_bytes(B, C, HW)
        ws = _workspace(y.device, nws)
        with torch.cuda.device(y.device):
            _launch(lib.sic_bottleneck_fwd(_ptr(y), _ptr(noise), _ptr(philox), _ptr(mu), _ptr(sigma), _ptr(nu), B, C, HW,
                                           quant_mode, lik_mode, layout, _ptr(y_tilde), _ptr(nll), _ptr(bits), _ptr(ws),
                                           ws.numel(), _stream()), "sic_bottleneck_fwd")
        if quant_mode == QUANT_NONE:
            y_tilde = y
        ctx.save_for_backward(y_tilde, sigma, nu, mu)
        ctx.cfg = (B, C, HW, quant_mode, lik_mode, layout)
        ctx.shapes = (sigma.shape, None if nu is None else nu.shape, None if mu is None else mu.shape)
        ctx.set_materialize_grads(False)
        return (y_tilde if quant_mode != QUANT_NONE else None), nll, bits

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_yt, g_nll, g_bits):
        lib = _lib.load()
        y_tilde, sigma, nu, mu = ctx.saved_tensors
        B, C, HW, quant_mode, lik_mode, layout = ctx.cfg
        need_y, need_s, need_n, need_m = ctx.needs_input_grad[:4]
        g_yt = None if g_yt is None else _require_cuda_f32(g_yt, "g_ytilde")
        g_nll = None if g_nll is None else _require_cuda_f32(g_nll, "g_nll")
        g_bits = None if g_bits is None else _require_cuda_f32(g_bits, "g_bits")
        dy = torch.empty_like(y_tilde) if need_y else None
        dsigma = torch.empty_like(sigma) if need_s else None
        dnu = torch.empty_like(nu) if (need_n and nu is not None) else None
        dmu = torch.empty_like(mu) if (need_m and mu is not None) else None
        ws = _workspace(y_tilde.device, lib.sic_bottleneck_workspace_bytes(B, C, HW))
        with torch.cuda.device(y_tilde.device):
            _launch(lib.sic_bottleneck_bwd(_ptr(y_tilde), _ptr(mu), _ptr(sigma), _ptr(nu), _ptr(g_nll), _ptr(g_bits), _ptr(g_yt),
                                           B, C, HW, quant_mode, lik_mode, layout, _ptr(dy), _ptr(dmu), _ptr(dsigma), _ptr(dnu),
                                           _ptr(ws), ws.numel(), _stream()), "sic_bottleneck_bwd")
        return dy, dsigma, dnu, dmu, None, None, None, None, None


def bottleneck(y: torch.Tensor, sigma: torch.Tensor, nu: Optional[torch.Tensor] = None, mu: Optional[torch.Tensor] = None,
               *, quant: str = "noise", lik: str = "density", noise: Optional[torch.Tensor] = None,
               want_nll: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
    """Fused quantise + likelihood + rate (K1).  Returns (y_tilde, nll, bits[B]).

    sigma / nu / mu: [B,C,1,1] (or [B,C]) -> broadcast layout; same shape as y -> spatial layout; for lik='gaussian'
    `sigma` is the per-channel log_sigma parameter [C].  quant: 'noise' | 'round' | 'none'; pass `noise=` to supply the
    uniform draw (parity mode) instead of the in-kernel Philox generator."""
    if quant not in _QUANT:
        raise ValueError(f"Unknown quant mode: {quant}")            # model.py:35
    if y.numel() == 0:                                               # empty batch / empty latent: nothing to launch
        _require_cuda_f32(y, "y")
        z = torch.zeros(y.shape[0], dtype=torch.float32, device=y.device)
        return y, (torch.empty_like(y) if want_nll else None), z + 0.0 * y.sum()
    q = _QUANT[quant]
    if quant == "noise" and noise is not None:
        q = QUANT_NOISE_TENSOR
    lik_mode = _LIK[lik]
    if lik == "gaussian":
        layout = PARAM_CHANNEL
    elif sigma.numel() == y.numel():
        layout = PARAM_SPATIAL
    elif sigma.numel() == y.shape[0] * y.shape[1]:
        layout = PARAM_BROADCAST
    else:
        raise _lib.SicError(f"sigma with {tuple(sigma.shape)} is neither per-(b,c) nor per-element for y {tuple(y.shape)}")
    y_tilde, nll, bits = _Bottleneck.apply(y, sigma, nu, mu, noise, q, lik_mode, layout, want_nll)
    return (y if y_tilde is None else y_tilde), nll, bits


def quantize(x: torch.Tensor, mode: str) -> torch.Tensor:
    """CompressionModel.quantize (model.py:27-35) on its own: kernel K1 with the likelihood output dropped.  'round' =
    half-to-even keeping -0.0 (zero gradient, like torch.round); 'noise' = x + U(-1/2, 1/2) from the in-kernel Philox stream
    (gradient 1).  CUDA float32 only."""
    if mode not in ("noise", "round"):
        raise ValueError(f"Unknown quant mode: {mode}")
    x = _require_cuda_f32(x, "x")
    if x.numel() == 0:
        return x
    return _Quantize.apply(x, mode)


class _Quantize(torch.autograd.Function):
    """Quantise-only form of K1.  The backward is the identity ('noise') or zero ('round', like torch.round): no kernel.  (Routing
    it through the likelihood kernel's backward cost 134 us per step in the overlapped schedule: with one row of 786 k elements the
    per-channel fold of that kernel's partials is a single thread's loop, r02v launch list.)"""

    @staticmethod
    @_amp_fwd
    def forward(ctx, x, mode: str):
        flat = x.reshape(1, 1, -1)
        zero = _zeros1.get(x.device)
        if zero is None:
            zero = _zeros1[x.device] = torch.zeros(1, dtype=torch.float32, device=x.device)
        with torch.no_grad():
            y_tilde, _, _ = bottleneck(flat, zero, quant=mode, lik="gaussian", want_nll=False)
        ctx.mode = mode
        return y_tilde.view(x.shape)

    @staticmethod
    @_amp_bwd
    def backward(ctx, g):
        return (g if ctx.mode == "noise" else torch.zeros_like(g)), None


_zeros1 = {}


# ----------------------------------------------------------------------------------------------------------------------
# K2
class _GDN(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, x, beta_param, gamma_weight, inverse: bool, bias=None):
        lib = _lib.load()
        x, cl = _dense_layout(x, "x")
        bias = None if bias is None else _require_cuda_f32(bias, "bias")
        back_to_cl = False
        if cl and x.shape[1] % 4 != 0:                # NHWC kernels walk channel quads; odd channel counts go through NCHW
            x, cl, back_to_cl = x.contiguous(), 0, True
        beta_param = _require_cuda_f32(beta_param, "beta")
        ctx.w_shape = gamma_weight.shape
        gamma_weight = _require_cuda_f32(gamma_weight.reshape(-1), "gamma_conv.weight")
        B, C = x.shape[0], x.shape[1]
        if beta_param.numel() != C or gamma_weight.numel() != C:
            raise _lib.SicError(f"GDN parameters have {beta_param.numel()}/{gamma_weight.numel()} entries for {C} channels")
        if bias is not None and bias.numel() != C:   # before the launch: the kernel indexes bias[c]
            raise _lib.SicError(f"bias has {bias.numel()} entries for {C} channels")
        HW = x.numel() // (B * C)
        y = torch.empty_like(x)                       # keeps x's memory format (NCHW or channels_last)
        with torch.cuda.device(x.device):
            _launch(lib.sic_gdn_fwd(_ptr(x), _ptr(bias), _ptr(beta_param), _ptr(gamma_weight), B, C, HW, int(inverse), cl, _ptr(y),
                                    _stream()), "sic_gdn_fwd")
        ctx.save_for_backward(x, beta_param, gamma_weight, bias)
        ctx.cfg = (B, C, HW, int(inverse), cl)
        return y.contiguous(memory_format=torch.channels_last) if back_to_cl else y

    @staticmethod
    @_amp_bwd
    def backward(ctx, g):
        lib = _lib.load()
        x, beta_param, gamma_weight, bias = ctx.saved_tensors
        B, C, HW, inverse, cl = ctx.cfg
        if cl:                                        # the gradient must be walked in x's memory order
            g = _dense_layout(g.contiguous(memory_format=torch.channels_last), "grad_output")[0]
        else:
            g = _require_cuda_f32(g, "grad_output")
        dx = torch.empty_like(x)
        dbeta = torch.empty_like(beta_param)
        dgamma = torch.empty_like(gamma_weight)
        dbias = torch.empty_like(bias) if bias is not None else None
        nws = lib.sic_gdn_bwd_workspace_bytes(B, C, HW)
        ws = _workspace(x.device, nws, "scratch")
        with torch.cuda.device(x.device):
            _lib.check(lib.sic_gdn_bwd(_ptr(x), _ptr(bias), _ptr(g), _ptr(beta_param), _ptr(gamma_weight), B, C, HW, inverse, cl,
                                       _ptr(dx), _ptr(dbias), _ptr(dbeta), _ptr(dgamma), _ptr(ws), ws.numel(), _stream()),
                       "sic_gdn_bwd")
        global launch_count
        launch_count += 2
        return dx, dbeta, dgamma.view(ctx.w_shape), None, dbias


def gdn(x: torch.Tensor, beta_param: torch.Tensor, gamma_weight: torch.Tensor, inverse: bool = False,
        bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Diagonal GDN/IGDN (K2), bit-exact with layers.py:19-27 on the same device arithmetic.  `bias`: the producing
    convolution's bias, folded in as GDN(x + bias) (same rounding as PyTorch's conv -> add_(bias) -> GDN)."""
    if x.numel() == 0:
        _require_cuda_f32(x, "x")
        return x + 0.0 * (beta_param.sum() + gamma_weight.sum())     # empty in, empty out (keeps the autograd graph connected)
    return _GDN.apply(x, beta_param, gamma_weight, inverse, bias)


class _GDNDense(torch.autograd.Function):
    """Dense-gamma GDN (G3), forward and backward on tcgen05 tensor cores.  Backward (SURVEY 8(a') G3) = three launches of
    libsic.so: pass 1 (s, h, direct, d(beta) partials), pass 2 (dx = direct + 2 x gamma^T h), pass 3 (d(gamma) = h^T x^2).
    The dense path is a capability the reference stores parameters for (layers.py:13) but never runs."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, x, beta_param, gamma_param, inverse: bool, variant=None):
        lib = _lib.load()
        if x.dim() != 4:
            raise _lib.SicError("dense GDN expects [B,C,H,W]")
        xc = _dense_layout(x.contiguous(memory_format=torch.channels_last), "x")[0]   # the kernel walks [positions, C]
        beta_param = _require_cuda_f32(beta_param, "beta")
        gamma_param = _require_cuda_f32(gamma_param, "gamma")
        B, C, H, W = x.shape
        if gamma_param.shape != (C, C) or beta_param.numel() != C:
            raise _lib.SicError(f"dense GDN parameters {tuple(beta_param.shape)}/{tuple(gamma_param.shape)} do not match C={C}")
        y = torch.empty_like(xc)
        with torch.cuda.device(x.device):
            if variant is None:
                rc = lib.sic_gdn_dense_fwd(_ptr(xc), _ptr(beta_param), _ptr(gamma_param), B * H * W, C, int(inverse), _ptr(y), _stream())
            else:
                rc = lib.sic_gdn_dense_fwd_variant(_ptr(xc), _ptr(beta_param), _ptr(gamma_param), B * H * W, C, int(inverse), _ptr(y),
                                                   int(variant), _stream())
            _launch(rc, "sic_gdn_dense_fwd")
        ctx.save_for_backward(xc, beta_param, gamma_param)
        ctx.inverse = bool(inverse)
        return y

    @staticmethod
    @_amp_bwd
    def backward(ctx, g):
        lib = _lib.load()
        xc, beta_param, gamma_param = ctx.saved_tensors
        B, C, H, W = xc.shape
        P = B * H * W
        gc = _dense_layout(g.contiguous(memory_format=torch.channels_last), "grad_output")[0]
        h, direct, dx = torch.empty_like(xc), torch.empty_like(xc), torch.empty_like(xc)
        rows = lib.sic_gdn_dense_bwd_part_rows(P, C)
        part = torch.empty((rows, C), dtype=torch.float32, device=xc.device)
        dgamma = torch.empty((C, C), dtype=torch.float32, device=xc.device)
        nws = lib.sic_gdn_dense_dgamma_workspace_bytes(P, C)
        ws = _workspace(xc.device, nws, "scratch")
        global launch_count
        with torch.cuda.device(xc.device):
            _lib.check(lib.sic_gdn_dense_bwd(_ptr(xc), _ptr(gc), _ptr(beta_param), _ptr(gamma_param), P, C, int(ctx.inverse), _ptr(h),
                                             _ptr(direct), _ptr(dx), _ptr(part), rows, _stream()), "sic_gdn_dense_bwd")
            launch_count += 2
            if ctx.needs_input_grad[2]:
                _lib.check(lib.sic_gdn_dense_dgamma(_ptr(xc), _ptr(h), P, C, _ptr(dgamma), _ptr(ws), ws.numel(), _stream()),
                           "sic_gdn_dense_dgamma")
                launch_count += 2
        dbeta = part.sum(0)                                          # fixed-order fold of the per-CTA partial rows
        return dx, dbeta * 2.0 * beta_param, (dgamma * 2.0 * gamma_param if ctx.needs_input_grad[2] else None), None, None


def gdn_dense(x: torch.Tensor, beta_param: torch.Tensor, gamma_param: torch.Tensor, inverse: bool = False,
              variant=None) -> torch.Tensor:
    """Dense-gamma GDN/IGDN on tcgen05 tensor cores (G3); returns a channels_last tensor.
    variant: None = the library default, _lib.DENSE_SERIAL or _lib.DENSE_PIPELINED to pick the kernel (include/sic.h)."""
    return _GDNDense.apply(x, beta_param, gamma_param, inverse, variant)


# ----------------------------------------------------------------------------------------------------------------------
# N2: first analysis layer (conv 3 -> C, 3x3 + bias + GDN) as one kernel
CONV0_CHANNELS = (32, 64, 96, 128, 192)


class _Conv0GDN(torch.autograd.Function):
    """layers.py:49-51 fused (csrc/conv0_gdn.cu).  The C x H x W intermediate is never stored: the backward recomputes it from the
    image on the tensor core and contracts dv with the im2col patches there too.  The image itself gets no gradient."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, x, weight, bias, beta_param, gamma_weight, want_v: bool):
        lib = _lib.load()
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise _lib.SicError("conv0_gdn: expected a CUDA tensor — this package has no CPU path (the CPU oracle is oracle/, test only)")
        if x.dim() != 4 or x.shape[1] != 3 or x.dtype != torch.float32:
            raise _lib.SicError(f"conv0_gdn: expected a float32 [B,3,H,W] image, got {tuple(x.shape)} {x.dtype}")
        if ctx.needs_input_grad[0]:
            raise _lib.SicError("conv0_gdn: the fused first layer does not return a gradient for the image")
        B, _, H, W = x.shape
        C = weight.shape[0]
        if tuple(weight.shape) != (C, 3, 3, 3) or C not in CONV0_CHANNELS:
            raise _lib.SicError(f"conv0_gdn: weight {tuple(weight.shape)} is not [C,3,3,3] with C in {CONV0_CHANNELS}")
        xh = x.permute(0, 2, 3, 1).contiguous()                                   # [B,H,W,3]; a view when x is channels_last
        wk = _require_cuda_f32(weight, "weight").permute(0, 2, 3, 1).contiguous()  # [C,kh,kw,cin]
        bias = None if bias is None else _require_cuda_f32(bias, "bias")
        beta_param = _require_cuda_f32(beta_param, "beta")
        ctx.w_shape = gamma_weight.shape
        gamma_weight = _require_cuda_f32(gamma_weight.reshape(-1), "gamma_conv.weight")
        if beta_param.numel() != C or gamma_weight.numel() != C or (bias is not None and bias.numel() != C):
            raise _lib.SicError(f"conv0_gdn: parameters do not match C={C}")
        y = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        v = torch.empty_like(y) if want_v else None
        with torch.cuda.device(x.device):
            _launch(lib.sic_conv0_gdn_fwd(_ptr(xh), _ptr(wk), _ptr(bias), _ptr(beta_param), _ptr(gamma_weight), B, H, W, C, _ptr(y), _ptr(v),
                                          _stream()), "sic_conv0_gdn_fwd")
        ctx.save_for_backward(xh, wk, bias, beta_param, gamma_weight)
        ctx.geom = (B, H, W, C)
        ctx.set_materialize_grads(False)
        if want_v:
            ctx.mark_non_differentiable(v)
        return y, v

    @staticmethod
    @_amp_bwd
    def backward(ctx, g, _gv):
        lib = _lib.load()
        xh, wk, bias, beta_param, gamma_weight = ctx.saved_tensors
        B, H, W, C = ctx.geom
        g = _dense_layout(g.contiguous(memory_format=torch.channels_last), "grad_output")[0]
        dev = xh.device
        dw = torch.empty((C, 3, 3, 3), dtype=torch.float32, device=dev)           # (kh, kw, cin) order
        dbias = torch.empty_like(bias) if bias is not None else None
        dbeta, dgamma = torch.empty_like(beta_param), torch.empty_like(gamma_weight)
        ws = _workspace(dev, lib.sic_conv0_gdn_bwd_workspace_bytes(B, H, W, C), "scratch")
        with torch.cuda.device(dev):
            _lib.check(lib.sic_conv0_gdn_bwd(_ptr(xh), _ptr(wk), _ptr(bias), _ptr(beta_param), _ptr(gamma_weight), _ptr(g), B, H, W, C,
                                             _ptr(dw), _ptr(dbias), _ptr(dbeta), _ptr(dgamma), _ptr(ws), ws.numel(), _stream()),
                       "sic_conv0_gdn_bwd")
        global launch_count
        launch_count += 2
        return None, dw.permute(0, 3, 1, 2), dbias, dbeta, dgamma.view(ctx.w_shape), None


def conv0_gdn(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], beta_param: torch.Tensor,
              gamma_weight: torch.Tensor, return_v: bool = False):
    """First analysis layer in one kernel (N2): GDN(conv2d(x, weight, bias, stride 1, padding 1)) for a [B,3,H,W] image and a
    [C,3,3,3] weight; returns a channels_last [B,C,H,W] tensor (and the pre-GDN activation with return_v=True, tests only).
    Tolerance-only training path (fp32-accurate tensor-core convolution, rsqrt GDN): see include/sic.h."""
    y, v = _Conv0GDN.apply(x, weight, bias, beta_param, gamma_weight, bool(return_v))
    return (y, v) if return_v else y


# ----------------------------------------------------------------------------------------------------------------------
# N2, synthesis side: the last layer ConvTranspose2d(N, 3, 5, 2, 2, output_padding=1) as GEMM + gather
class _Col2ImRGB(torch.autograd.Function):
    """The scatter half of layers.py:96-98 (`deconv(N, 3)`) as a deterministic gather (csrc/deconv_rgb.cu): D [B,80,H,W] channels-last
    (75 tap columns (kh, kw, co) + 5 zero columns per position) -> x_hat [B,3,2H,2W] channels-last; backward = the adjoint gather."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, D, bias):
        lib = _lib.load()
        if D.dim() != 4 or D.shape[1] != 80:
            raise _lib.SicError("col2im_rgb expects D [B,80,H,W]")
        Dc = _dense_layout(D.contiguous(memory_format=torch.channels_last), "D")[0]
        bias = None if bias is None else _require_cuda_f32(bias, "bias")
        B, _, H, W = Dc.shape
        out = torch.empty((B, 3, 2 * H, 2 * W), dtype=torch.float32, device=D.device, memory_format=torch.channels_last)
        with torch.cuda.device(D.device):
            _launch(lib.sic_deconv_rgb_col2im(_ptr(Dc), _ptr(bias), B, H, W, _ptr(out), _stream()), "sic_deconv_rgb_col2im")
        ctx.geom = (B, H, W, bias is not None)
        return out

    @staticmethod
    @_amp_bwd
    def backward(ctx, g):
        lib = _lib.load()
        B, H, W, has_bias = ctx.geom
        g = _dense_layout(g.contiguous(memory_format=torch.channels_last), "grad_output")[0]
        dD = None
        if ctx.needs_input_grad[0]:
            dD = torch.empty((B, 80, H, W), dtype=torch.float32, device=g.device, memory_format=torch.channels_last)
            with torch.cuda.device(g.device):
                _launch(lib.sic_deconv_rgb_im2col(_ptr(g), B, H, W, _ptr(dD), _stream()), "sic_deconv_rgb_im2col")
        dbias = channel_sum(g) if (has_bias and ctx.needs_input_grad[1]) else None
        return dD, dbias


def deconv_rgb(a: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """conv_transpose2d(a, weight [N,3,5,5], bias, stride 2, padding 2, output_padding 1) as  D = 1x1 convolution N -> 75 (+5 zero
    columns; cuDNN's sm_100 tensor-op kernels, forward / dgrad / wgrad through autograd)  followed by the deterministic col2im
    gather; returns a channels_last [B,3,2H,2W] tensor.  cuDNN runs the 3-band transposed convolution itself on legacy
    non-tensor-core engines (238 us forward, 438 us backward per cfg2 step).  Same precision policy as the layer it replaces
    (torch.backends.cudnn.allow_tf32 governs both).  Opt-in training path (layers.FAST_LAST_LAYER); eval/compress keep the
    cuDNN layer, whose x_hat is pinned bit for bit."""
    if not (isinstance(a, torch.Tensor) and a.is_cuda):
        raise _lib.SicError("deconv_rgb: expected a CUDA tensor — this package has no CPU path (the CPU oracle is oracle/, test only)")
    N = a.shape[1]
    if a.dim() != 4 or tuple(weight.shape) != (N, 3, 5, 5):
        raise _lib.SicError(f"deconv_rgb: weight {tuple(weight.shape)} is not [{N},3,5,5] for activations {tuple(a.shape)}")
    wm = weight.permute(2, 3, 1, 0).reshape(75, N)                              # row m = (kh, kw, co)
    wm = torch.nn.functional.pad(wm, (0, 0, 0, 5)).view(80, N, 1, 1)
    D = torch.nn.functional.conv2d(a.contiguous(memory_format=torch.channels_last), wm)
    return _Col2ImRGB.apply(D, bias)


# ----------------------------------------------------------------------------------------------------------------------
# N4: tail of the hyper-synthesis transform
class _HyperTail(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, t, w1s, b1s, w2s, b2s, w1n, b1n, w2n, b2n, min_nu: float, max_nu: float):
        lib = _lib.load()
        t, cl = _dense_layout(t, "t")
        if t.dim() != 4:
            raise _lib.SicError("hyper_tail expects the [B,N,h,w] output of h_s.h_s")
        B, N, H, W = t.shape
        M = w2s.shape[0]
        ws = [_require_cuda_f32(w, "mlp parameter") for w in (w1s, b1s, w2s, b2s, w1n, b1n, w2n, b2n)]
        want = [N * N, N, M * N, M] * 2
        if [w.numel() for w in ws] != want or w2n.shape[0] != M:
            raise _lib.SicError(f"hyper_tail: MLP parameters do not match N={N}, M={M}")
        sigma = torch.empty((B, M, 1, 1), dtype=torch.float32, device=t.device)
        nu = torch.empty((B, M, 1, 1), dtype=torch.float32, device=t.device)
        need = any(ctx.needs_input_grad[:9])
        save = torch.empty(lib.sic_hyper_tail_save_floats(B, N, M), dtype=torch.float32, device=t.device) if need else None
        with torch.cuda.device(t.device):
            _launch(lib.sic_hyper_tail_fwd(_ptr(t), B, N, M, H * W, cl, *[_ptr(w) for w in ws], float(min_nu), float(max_nu), _ptr(sigma),
                                           _ptr(nu), _ptr(save), _stream()), "sic_hyper_tail_fwd")
        if need:
            ctx.save_for_backward(sigma, save, *ws)
        ctx.cfg = (B, N, M, H, W, cl, float(min_nu), float(max_nu), t.shape, t.stride())
        ctx.shapes = [w.shape for w in (w1s, b1s, w2s, b2s, w1n, b1n, w2n, b2n)]
        ctx.set_materialize_grads(False)
        return sigma, nu

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_sigma, g_nu):
        lib = _lib.load()
        sigma, save, w1s, b1s, w2s, b2s, w1n, b1n, w2n, b2n = ctx.saved_tensors
        B, N, M, H, W, cl, min_nu, max_nu, t_shape, t_stride = ctx.cfg
        g_sigma = None if g_sigma is None else _require_cuda_f32(g_sigma, "g_sigma")
        g_nu = None if g_nu is None else _require_cuda_f32(g_nu, "g_nu")
        dev = sigma.device
        dt = torch.empty_strided(t_shape, t_stride, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        grads = [torch.empty(shp, dtype=torch.float32, device=dev) for shp in ctx.shapes]
        scratch = torch.empty(lib.sic_hyper_tail_scratch_floats(B, N, M), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.sic_hyper_tail_bwd(_ptr(g_sigma), _ptr(g_nu), _ptr(sigma), _ptr(save), B, N, M, H * W, cl, _ptr(w1s), _ptr(w2s),
                                              _ptr(w1n), _ptr(w2n), min_nu, max_nu, _ptr(dt), *[_ptr(g) for g in grads], _ptr(scratch),
                                              _stream()), "sic_hyper_tail_bwd")
        global launch_count
        launch_count += 2
        return (dt, *grads, None, None)


def hyper_tail(t: torch.Tensor, mlp_sigma, mlp_nu, min_nu: float, max_nu: float):
    """N4: pool -> mlp_sigma / mlp_nu -> exp / clamp in one launch (layers.py:141-152 + model.py:54-55).  `t`: output of h_s.h_s;
    mlp_*: the reference's nn.Sequential(Conv2d(N,N,1), ReLU, Conv2d(N,M,1)).  Returns (sigma, nu) as [B,M,1,1] (K1's layout)."""
    return _HyperTail.apply(t, mlp_sigma[0].weight, mlp_sigma[0].bias, mlp_sigma[2].weight, mlp_sigma[2].bias,
                            mlp_nu[0].weight, mlp_nu[0].bias, mlp_nu[2].weight, mlp_nu[2].bias, float(min_nu), float(max_nu))


# ----------------------------------------------------------------------------------------------------------------------
# N3: SSIM statistics of one MS-SSIM scale
class _SSIMStats(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, x, y, c1: float, c2: float, want_pool: bool):
        lib = _lib.load()
        x = _require_cuda_f32(x, "x")
        y = _require_cuda_f32(y, "y")
        if x.shape != y.shape or x.dim() != 4:
            raise _lib.SicError("ssim_stats expects two [B,C,H,W] tensors of the same shape")
        B, C, H, W = x.shape
        if H < 11 or W < 11:
            raise ValueError("Kernel size can't be greater than actual input size.")
        if want_pool and ((H | W) & 1):
            raise _lib.SicError("ssim_stats: pooled outputs need even H and W")
        planes, tiles = B * C, lib.sic_ssim_tiles(H, W)
        part = torch.empty((2, planes, tiles), dtype=torch.float32, device=x.device)
        need = ctx.needs_input_grad[0]
        maps = torch.empty((5, planes, H - 10, W - 10), dtype=torch.float32, device=x.device) if need else None
        xp = torch.empty((B, C, H // 2, W // 2), dtype=torch.float32, device=x.device) if want_pool else None
        yp = torch.empty_like(xp) if want_pool else None
        with torch.cuda.device(x.device):
            _launch(lib.sic_ssim_fwd_pool(_ptr(x), _ptr(y), planes, H, W, c1, c2, _ptr(part[0]), _ptr(part[1]), _ptr(maps), _ptr(xp), _ptr(yp),
                                          _stream()), "sic_ssim_fwd")
        means = part.sum(dim=2) * (1.0 / ((H - 10) * (W - 10)))             # fixed-order fold of the per-tile partials
        if need:
            ctx.save_for_backward(x, y, maps)
        ctx.set_materialize_grads(False)
        if want_pool:
            ctx.mark_non_differentiable(yp)
        return means[0].view(B, C), means[1].view(B, C), xp, yp

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_ss, g_cs, g_xp, _g_yp):
        lib = _lib.load()
        x, y, maps = ctx.saved_tensors
        B, C, H, W = x.shape
        g_ss = None if g_ss is None else _require_cuda_f32(g_ss, "g_ss")
        g_cs = None if g_cs is None else _require_cuda_f32(g_cs, "g_cs")
        g_xp = None if g_xp is None else _require_cuda_f32(g_xp, "g_x_pool")
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _launch(lib.sic_ssim_bwd_pool(_ptr(x), _ptr(y), _ptr(maps), _ptr(g_ss), _ptr(g_cs), _ptr(g_xp), B * C, H, W, _ptr(dx), _stream()),
                    "sic_ssim_bwd")
        return dx, None, None, None, None


def ssim_stats(x: torch.Tensor, y: torch.Tensor, c1: float = 0.01 ** 2, c2: float = 0.03 ** 2, pool: bool = False):
    """Per-(batch, channel) means of the SSIM map and of its contrast-structure factor for one scale: (ss [B,C], cs [B,C]).
    Differentiable w.r.t. x only (y is the target).  pool=True (even H, W): also returns the next MS-SSIM scale's inputs, the 2x2
    average-pooled x and y, written by the same kernel - (ss, cs, x_pool, y_pool); the gradient that later arrives at x_pool is added
    inside this scale's backward kernel."""
    ss, cs, xp, yp = _SSIMStats.apply(x, y, float(c1), float(c2), bool(pool))
    return (ss, cs, xp, yp) if pool else (ss, cs)


class _MSSSIM(torch.autograd.Function):
    """All scales of the MS-SSIM value in L + 1 launches forward (one statistics kernel per scale, each also writing the next
    scale's pooled inputs, then the combination kernel) and L launches backward (coarsest scale first; each hands its dX to the next
    finer one as the pooled gradient).  No torch op in between: the clamp of the reconstruction, the layout of both images
    (planar or channels-last), the fold of the partial sums, relu / pow / product / mean and the incoming gradient's scale all
    live in the kernels."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, x, y, weights, normalize: bool, c1: float, c2: float, clamp01: bool):
        lib = _lib.load()
        x, x_cl = _dense_layout(x, "x")
        y, y_cl = _dense_layout(y, "y")
        weights = _require_cuda_f32(weights, "scale_weights")
        if x.shape != y.shape or x.dim() != 4:
            raise _lib.SicError("multi_scale_ssim expects two [B,C,H,W] tensors of the same shape")
        B, C, H, W = x.shape
        L, planes = weights.numel(), B * C
        sizes = [(H >> l, W >> l) for l in range(L)]
        if any((h | w) & 1 for h, w in sizes[:-1]) or sizes[-1][0] < 11 or sizes[-1][1] < 11:
            raise _lib.SicError("multi_scale_ssim (fused): every pooled scale needs even H and W and the last one at least 11x11")
        tiles = [int(lib.sic_ssim_tiles(h, w)) for h, w in sizes]
        offs, total = [], 0
        for t in tiles:
            offs.append(total)
            total += 2 * planes * t
        need = ctx.needs_input_grad[0]
        part = torch.empty(total, dtype=torch.float32, device=x.device)
        out = torch.empty((), dtype=torch.float32, device=x.device)
        coef = torch.empty((L, 2, planes), dtype=torch.float32, device=x.device) if need else None
        saved, xs, ys = [], x, y
        with torch.cuda.device(x.device):
            for l, (h, w) in enumerate(sizes):
                first, last = l == 0, l == L - 1
                maps = torch.empty((5, planes, h - 10, w - 10), dtype=torch.float32, device=x.device) if need else None
                xp = None if last else torch.empty((B, C, h // 2, w // 2), dtype=torch.float32, device=x.device)
                yp = None if last else torch.empty_like(xp)
                ps = part[offs[l]:offs[l] + planes * tiles[l]]
                pc = part[offs[l] + planes * tiles[l]:offs[l] + 2 * planes * tiles[l]]
                _launch(lib.sic_ssim_fwd_ex(_ptr(xs), _ptr(ys), planes, h, w, C if (first and x_cl) else 0, C if (first and y_cl) else 0,
                                            int(bool(clamp01) and first), c1, c2, _ptr(ps), _ptr(pc), _ptr(maps), _ptr(xp), _ptr(yp),
                                            _stream()), "sic_ssim_fwd_ex")
                saved += [xs, ys, maps]
                xs, ys = xp, yp
            _launch(lib.sic_msssim_combine(_ptr(part), (ctypes.c_long * L)(*offs), (ctypes.c_int * L)(*tiles),
                                           (ctypes.c_long * L)(*[(h - 10) * (w - 10) for h, w in sizes]), L, planes, _ptr(weights),
                                           int(bool(normalize)), _ptr(out), _ptr(coef), _stream()), "sic_msssim_combine")
        if need:
            ctx.save_for_backward(coef, *saved)
            ctx.geom = (sizes, planes, C, x_cl, y_cl, bool(clamp01))
        return out

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_out):
        lib = _lib.load()
        coef, *saved = ctx.saved_tensors
        sizes, planes, C, x_cl, y_cl, clamp01 = ctx.geom
        L = len(sizes)
        g_out = _require_cuda_f32(g_out, "grad of the MS-SSIM value")
        g_pool = None
        with torch.cuda.device(g_out.device):
            for l in range(L - 1, -1, -1):
                h, w = sizes[l]
                xs, ys, maps = saved[3 * l], saved[3 * l + 1], saved[3 * l + 2]
                first, last = l == 0, l == L - 1
                dx = torch.empty_like(xs)                      # level 0: the layout of x (dense NCHW or channels-last) is kept
                _launch(lib.sic_ssim_bwd_ex(_ptr(xs), _ptr(ys), _ptr(maps), _ptr(coef[l, 0]) if last else None,
                                            None if last else _ptr(coef[l, 1]), _ptr(g_out), _ptr(g_pool), planes, h, w,
                                            C if (first and x_cl) else 0, C if (first and y_cl) else 0, int(clamp01 and first), _ptr(dx),
                                            _stream()), "sic_ssim_bwd_ex")
                g_pool = dx
        return g_pool, None, None, None, None, None, None


def msssim_fused(x: torch.Tensor, y: torch.Tensor, scale_weights: torch.Tensor, normalize: bool = True, c1: float = 0.01 ** 2,
                 c2: float = 0.03 ** 2, clamp01: bool = False) -> torch.Tensor:
    """MS-SSIM value (0-dim) of x against the target y over len(scale_weights) dyadic scales, differentiable w.r.t. x only; every scale
    that is pooled must have even H and W (losses.multi_scale_ssim falls back to the per-scale kernels + torch pooling otherwise).
    scale_weights: CUDA float32 [L]; normalize: divide them by their sum (inside the kernel).  clamp01: x is clamped to [0, 1] first."""
    return _MSSSIM.apply(x, y, scale_weights, bool(normalize), float(c1), float(c2), bool(clamp01))


# ----------------------------------------------------------------------------------------------------------------------
# K4 / K3 / E1
def quantize_indices(q: torch.Tensor, do_round: bool = False, tail: int = 10, want_symbols: bool = True):
    """Per-patch support and symbols (K4): returns (sym int32 like q, mins int32 [B], maxs int32 [B]) on the device."""
    lib = _lib.load()
    q = _require_cuda_f32(q, "q")
    B = q.shape[0]
    if q.numel() == 0:
        raise _lib.SicError("quantize_indices: empty latent (no patches or no elements) has no support")
    n_per = q.numel() // B
    sym = torch.empty(q.shape, dtype=torch.int32, device=q.device) if want_symbols else None
    mins = torch.empty(B, dtype=torch.int32, device=q.device)
    maxs = torch.empty(B, dtype=torch.int32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(lib.sic_quantize_indices(_ptr(q), B, n_per, int(do_round), int(tail), _ptr(sym), _ptr(mins), _ptr(maxs),
                                            _stream()), "sic_quantize_indices")
    global launch_count
    launch_count += 4 if want_symbols else 3
    return sym, mins, maxs


def build_cdf_tables(kind: str, sigma: torch.Tensor, nu: Optional[torch.Tensor], n_patches: int, mins: torch.Tensor,
                     maxs: torch.Tensor, stride: int, channels: Optional[int] = None) -> torch.Tensor:
    """uint16 CDF tables [n_rows, stride] (K3).  kind 'gaussian': sigma = log_sigma[C], one row per (patch, channel);
    kind 'studentt': sigma/nu raw, n_rows = sigma.numel(), rows split evenly over the patches."""
    lib = _lib.load()
    sigma = _require_cuda_f32(sigma, "sigma")
    if kind == "gaussian":
        C = sigma.numel()
        n_rows, rows_per_patch, k = n_patches * C, C, 0
    elif kind == "studentt":
        nu = _require_cuda_f32(nu, "nu")
        n_rows = sigma.numel()
        if n_rows % n_patches or nu.numel() != n_rows:
            raise _lib.SicError("sigma/nu row count must be a multiple of the patch count")
        rows_per_patch, C, k = n_rows // n_patches, channels or (n_rows // n_patches), 1
    else:
        raise ValueError("kind must be 'gaussian' or 'studentt'")
    if mins.dtype != torch.int32 or maxs.dtype != torch.int32 or not mins.is_cuda:
        raise _lib.SicError("mins/maxs must be CUDA int32 tensors")
    out = torch.empty((n_rows, stride), dtype=torch.uint16, device=sigma.device)
    with torch.cuda.device(sigma.device):
        _launch(lib.sic_build_cdf_tables(k, _ptr(sigma), _ptr(nu), n_rows, rows_per_patch, C, _ptr(mins.contiguous()),
                                         _ptr(maxs.contiguous()), int(stride), _ptr(out), _stream()), "sic_build_cdf_tables")
    return out


def rans_encode(sym: np.ndarray, tables: np.ndarray, L: int, sym_per_row: int) -> bytes:
    """Host rANS encoder (E1, SIC-RANS-1).  sym int32 [n], tables uint16 [rows, stride]."""
    lib = _lib.load()
    sym = np.ascontiguousarray(sym, np.int32).ravel()
    tables = np.ascontiguousarray(tables, np.uint16)
    cap = 128 + 2 * sym.size + 16
    out = np.empty(cap, np.uint8)
    n = lib.sic_rans_encode_host(sym.ctypes.data, sym.size, tables.ctypes.data, tables.shape[-1], int(L), int(sym_per_row),
                                 out.ctypes.data, cap)
    if n < 0:
        _lib.check(int(n), "sic_rans_encode_host")
    return out[:n].tobytes()


def rans_decode(data: bytes, n: int, tables: np.ndarray, L: int, sym_per_row: int) -> np.ndarray:
    lib = _lib.load()
    tables = np.ascontiguousarray(tables, np.uint16)
    buf = np.frombuffer(data, np.uint8)
    sym = np.empty(int(n), np.int32)
    _lib.check(lib.sic_rans_decode_host(buf.ctypes.data, buf.size, int(n), tables.ctypes.data, tables.shape[-1], int(L),
                                        int(sym_per_row), sym.ctypes.data), "sic_rans_decode_host")
    return sym


def rans_encode_device(sym: torch.Tensor, tables: torch.Tensor, Ls: torch.Tensor, sym_per_row: int, rows_per_stream: int):
    """GPU rANS encoder (N1): sym int32 [S, n], tables uint16 [S*rows_per_stream, stride], Ls int32 [S] — all on the device.
    Returns (bytes uint8 [S, cap], nbytes int32 [S]) on the device; one warp per stream, byte-identical to the host coder."""
    lib = _lib.load()
    if not (sym.is_cuda and sym.dtype == torch.int32 and tables.dtype == torch.uint16 and Ls.dtype == torch.int32):
        raise _lib.SicError("rans_encode_device: expected CUDA int32 symbols, uint16 tables, int32 Ls")
    S = sym.shape[0]
    n = sym.numel() // S
    sym, tables, Ls = sym.contiguous(), tables.contiguous(), Ls.contiguous()
    cap = (128 + 2 * n + 3) // 4 * 4
    out = torch.empty((S, cap), dtype=torch.uint8, device=sym.device)
    nbytes = torch.empty(S, dtype=torch.int32, device=sym.device)
    ws = _workspace(sym.device, lib.sic_rans_encode_workspace_bytes(S, n), "scratch")
    with torch.cuda.device(sym.device):
        _lib.check(lib.sic_rans_encode_ws(_ptr(sym), _ptr(tables), _ptr(Ls), S, n, int(sym_per_row), int(rows_per_stream), tables.shape[-1],
                                          _ptr(out), cap, _ptr(nbytes), _ptr(ws), ws.numel(), _stream()), "sic_rans_encode_ws")
    global launch_count
    launch_count += 2
    return out, nbytes


def rans_decode_device(data: torch.Tensor, nbytes: torch.Tensor, tables: torch.Tensor, Ls: torch.Tensor, n: int, sym_per_row: int,
                       rows_per_stream: int):
    """GPU rANS decoder (N1): data uint8 [S, cap] (cap % 4 == 0), nbytes int32 [S]; returns (sym int32 [S, n], status int32 [S])."""
    lib = _lib.load()
    S, cap = data.shape
    if not (data.is_cuda and data.dtype == torch.uint8 and nbytes.dtype == torch.int32 and cap % 4 == 0):
        raise _lib.SicError("rans_decode_device: expected CUDA uint8 [S, cap] with cap % 4 == 0 and int32 nbytes")
    data, tables, Ls, nbytes = data.contiguous(), tables.contiguous(), Ls.contiguous(), nbytes.contiguous()
    sym = torch.empty((S, n), dtype=torch.int32, device=data.device)
    status = torch.empty(S, dtype=torch.int32, device=data.device)
    with torch.cuda.device(data.device):
        _launch(lib.sic_rans_decode(_ptr(data), _ptr(nbytes), _ptr(tables), _ptr(Ls), S, int(n), int(sym_per_row), int(rows_per_stream),
                                    tables.shape[-1], cap, _ptr(sym), _ptr(status), _stream()), "sic_rans_decode")
    return sym, status


# ----------------------------------------------------------------------------------------------------------------------
# (e) training-step tail on the flat buffers
def clip_adam_workspace(n: int, device: torch.device) -> torch.Tensor:
    lib = _lib.load()
    return torch.empty(max(int(lib.sic_clip_adam_workspace_bytes(int(n))), 8) // 8, dtype=torch.float64, device=device)


def clip_adam_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: torch.Tensor,
                   norm_out: Optional[torch.Tensor], workspace: torch.Tensor, *, inv_world: float, clip: float, lr: float,
                   betas: Tuple[float, float], eps: float, weight_decay: float) -> None:
    """Global-norm clip (train.py:200-202) + Adam (train.py:182-183) on flat float32 buffers, in place, two launches; `step` is the
    device-side update counter (float32 scalar, incremented by the call), `norm_out` receives the pre-clip norm of the mean gradient."""
    global launch_count
    lib = _lib.load()
    for name, t in (("param", param), ("grad", grad), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 1 and t.is_contiguous()):
            raise _lib.SicError(f"clip_adam_step: {name} must be a flat contiguous CUDA float32 tensor")
        if t.numel() != param.numel():
            raise _lib.SicError(f"clip_adam_step: {name} has {t.numel()} elements, param has {param.numel()}")
    if step.dtype != torch.float32 or step.numel() != 1 or not step.is_cuda:
        raise _lib.SicError("clip_adam_step: step must be a CUDA float32 scalar")
    with torch.cuda.device(param.device):
        _launch(lib.sic_clip_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), _ptr(step),
                                       float(inv_world), float(clip or 0.0), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                       float(weight_decay), _ptr(norm_out), _ptr(workspace), workspace.numel() * 8, _stream()),
                "sic_clip_adam_step")
    launch_count += 1


class _RDLossTail(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, bits_y, bits_z, dist, pixels: int, lambda_rd: float, similarity: bool):
        lib = _lib.load()
        bits_y = _require_cuda_f32(bits_y, "bits_y").reshape(-1)
        bits_z = _require_cuda_f32(bits_z, "bits_z").reshape(-1)
        dist = _require_cuda_f32(dist, "distortion value")
        if dist.numel() != 1:
            raise _lib.SicError("rd_loss_tail: the distortion must be a scalar")
        dev = bits_y.device
        loss, R, D, keep = (torch.empty((), dtype=torch.float32, device=dev) for _ in range(4))
        with torch.cuda.device(dev):
            _launch(lib.sic_rd_loss_fwd(_ptr(bits_y), bits_y.numel(), _ptr(bits_z), bits_z.numel(), _ptr(dist), int(similarity), int(pixels),
                                        float(lambda_rd), _ptr(loss), _ptr(R), _ptr(D), _ptr(keep), _stream()), "sic_rd_loss_fwd")
        ctx.save_for_backward(keep)
        ctx.geom = (bits_y.numel(), bits_z.numel(), int(pixels), float(lambda_rd), bool(similarity), dist.shape)
        ctx.mark_non_differentiable(R, D)
        return loss, R, D

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_loss, _g_R, _g_D):
        lib = _lib.load()
        (keep,) = ctx.saved_tensors
        ny, nz, pixels, lambda_rd, similarity, dshape = ctx.geom
        g_loss = _require_cuda_f32(g_loss, "g_loss")
        gy = torch.empty(ny, dtype=torch.float32, device=keep.device)
        gz = torch.empty(nz, dtype=torch.float32, device=keep.device)
        gd = torch.empty(dshape, dtype=torch.float32, device=keep.device)
        with torch.cuda.device(keep.device):
            _launch(lib.sic_rd_loss_bwd(_ptr(g_loss), _ptr(keep), pixels, lambda_rd, int(similarity), ny, nz, _ptr(gy), _ptr(gz), _ptr(gd),
                                        _stream()), "sic_rd_loss_bwd")
        return gy, gz, gd, None, None, None


def rd_loss_tail(bits_y: torch.Tensor, bits_z: torch.Tensor, dist: torch.Tensor, pixels: int, lambda_rd: float, similarity: bool):
    """(loss, R, D) of model.py:75-107 from the per-patch bit counts of y and z and the distortion value (the MS-SSIM when
    `similarity`, else the MSE): R = max((sum bits_y + sum bits_z) / pixels, 0), D = 1 - dist | dist, loss = lambda D + R.
    One launch forward, one backward; R and D carry no gradient (the reference returns them detached)."""
    return _RDLossTail.apply(bits_y, bits_z, dist, int(pixels), float(lambda_rd), bool(similarity))


def pack_flat(tensors, dst_offsets, dst: torch.Tensor) -> None:
    """Copy every (flat, contiguous, CUDA float32) tensor of `tensors` to dst[dst_offsets[t] : dst_offsets[t] + numel] in one launch per
    128 tensors (the gradient pack of FlatTrainer; replaces torch.cat(..., out=slice))."""
    global launch_count
    lib = _lib.load()
    n = len(tensors)
    if n == 0:
        return
    for t, off in zip(tensors, dst_offsets):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == dst.device):
            raise _lib.SicError("pack_flat: expected contiguous CUDA float32 tensors on the destination's device")
        if off < 0 or off + t.numel() > dst.numel():
            raise _lib.SicError(f"pack_flat: slice [{off}, {off + t.numel()}) outside the destination of {dst.numel()} elements")
    srcs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
    numels = (ctypes.c_long * n)(*[t.numel() for t in tensors])
    offs = (ctypes.c_long * n)(*[int(o) for o in dst_offsets])
    with torch.cuda.device(dst.device):
        _launch(lib.sic_pack_flat(srcs, numels, offs, n, _ptr(dst), _stream()), "sic_pack_flat")
    launch_count += (n - 1) // 128


# ----------------------------------------------------------------------------------------------------------------------
# bias add (+ ReLU) after a bias-free convolution, channels-last activations
class _BiasAct(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, t, bias, relu: bool):
        lib = _lib.load()
        bias = _require_cuda_f32(bias, "bias")
        B, C, H, W = t.shape
        with torch.cuda.device(t.device):
            _launch(lib.sic_bias_act_fwd(_ptr(t), _ptr(bias), B * H * W, C, int(relu), _stream()), "sic_bias_act_fwd")
        ctx.mark_dirty(t)
        ctx.relu = bool(relu)
        if relu:
            ctx.save_for_backward(t)
        return t

    @staticmethod
    @_amp_bwd
    def backward(ctx, g):
        global launch_count
        lib = _lib.load()
        g = _require_cuda_f32_cl(g, "grad_output")
        B, C, H, W = g.shape
        P = B * H * W
        y = ctx.saved_tensors[0] if ctx.relu else None
        dt = torch.empty_like(g) if ctx.relu else None
        dbias = torch.empty(C, dtype=torch.float32, device=g.device)
        ws = _workspace(g.device, int(lib.sic_bias_grad_workspace_bytes(P, C)), kind="bias_grad")
        with torch.cuda.device(g.device):
            _launch(lib.sic_bias_act_bwd(_ptr(g), _ptr(y), P, C, _ptr(dt), _ptr(dbias), _ptr(ws), ws.numel(), _stream()), "sic_bias_act_bwd")
        launch_count += 1
        return (dt if ctx.relu else g), dbias, None


def _is_channels_last_dense(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


def _require_cuda_f32_cl(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or t.dim() != 4:
        raise _lib.SicError(f"{name}: expected a 4-D CUDA float32 tensor")
    return t if _is_channels_last_dense(t) else t.contiguous(memory_format=torch.channels_last)


def bias_act_supported(t: torch.Tensor, bias: Optional[torch.Tensor]) -> bool:
    """True when bias_act can take `t` as it is: a dense channels-last CUDA float32 activation with a float32 bias of at most 1024 channels."""
    return (bias is not None and isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and bias.dtype == torch.float32
            and _is_channels_last_dense(t) and t.shape[1] <= 1024 and t.numel() > 0)


def bias_act(t: torch.Tensor, bias: torch.Tensor, relu: bool = False) -> torch.Tensor:
    """t + bias[None, :, None, None] (then ReLU) IN PLACE on the channels-last output `t` of a bias-free convolution - PyTorch's own
    add_(bias) / relu_ in one launch, same values - with d(bias) (and the ReLU mask) of the backward as one pass + a fold instead of
    threshold_backward + a reduction over (B, H, W).  See bias_act_supported()."""
    if not bias_act_supported(t, bias):
        raise _lib.SicError("bias_act: expects a dense channels-last CUDA float32 activation and a float32 bias (no CPU path)")
    return _BiasAct.apply(t, bias, bool(relu))


def channel_sum(g: torch.Tensor) -> torch.Tensor:
    """sum over (B, H, W) of a channels-last [B,C,H,W] CUDA float32 tensor -> [C], deterministic (the d(bias) reduction)."""
    global launch_count
    lib = _lib.load()
    g = _require_cuda_f32_cl(g, "g")
    B, C, H, W = g.shape
    P = B * H * W
    out = torch.empty(C, dtype=torch.float32, device=g.device)
    ws = _workspace(g.device, int(lib.sic_bias_grad_workspace_bytes(P, C)), kind="bias_grad")
    with torch.cuda.device(g.device):
        _launch(lib.sic_bias_act_bwd(_ptr(g), None, P, C, None, _ptr(out), _ptr(ws), ws.numel(), _stream()), "sic_bias_act_bwd")
    launch_count += 1
    return out
