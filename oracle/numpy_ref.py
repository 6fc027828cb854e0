"""numpy restatement of the reference's bottleneck + GDN arithmetic (test infrastructure).

fp32 functions replay the reference's operation ORDER with IEEE-754 single
precision numpy ops (numpy's + - * / sqrt are correctly rounded, like CUDA's
`__f*_rn` and unlike torch-CPU's MKL-VML sqrt, see DESIGN.md "oracle notes").
`*_f64` functions are the float64 truth used to separate "we differ from the
reference" from "the reference's own fp32 noise".

All file:line citations are relative to /root/reference/code/modelv2.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import special as sp

F32 = np.float32
LOG2E = 1.0 / math.log(2.0)          # distributions.py:6 (python double)
LOG2E_F32 = F32(LOG2E)               # what the fp32 tensor multiply sees
SIGMA_MIN, SIGMA_MAX = 1e-3, 1e3     # distributions.py:23,43
NU_MIN, NU_MAX = 2.0, 100.0          # distributions.py:24
REPARAM_OFFSET = F32(2.0 ** -18)     # layers.py:8


# --------------------------------------------------------------------------- quantize
def quantize_noise(x: np.ndarray, noise: np.ndarray) -> np.ndarray:
    """model.py:28-31 with the uniform(-.5,.5) draw supplied by the caller."""
    return (x.astype(F32) + noise.astype(F32)).astype(F32)


def quantize_round(x: np.ndarray) -> np.ndarray:
    """model.py:32-33: torch.round == round-half-to-even, keeps -0.0."""
    return np.rint(x.astype(F32)).astype(F32)


# --------------------------------------------------------------------------- Student-t density (L1)
def studentt_nll_f32(x, sigma, nu):
    """distributions.py:20-31 in float32, same op order (lgamma via scipy in fp32)."""
    x = np.asarray(x, F32)
    sigma = np.clip(np.asarray(sigma, F32), F32(SIGMA_MIN), F32(SIGMA_MAX))
    nu = np.clip(np.asarray(nu, F32), F32(NU_MIN), F32(NU_MAX))
    half = F32(0.5)
    logC = (sp.gammaln(((nu + F32(1.0)) / F32(2.0)).astype(F32)).astype(F32)
            - sp.gammaln((nu / F32(2.0)).astype(F32)).astype(F32)
            - half * np.log((nu * F32(math.pi)).astype(F32))
            - np.log(sigma)).astype(F32)
    quad = ((x / sigma) ** 2).astype(F32)
    logp = logC - ((nu + F32(1.0)) / F32(2.0)) * np.log1p((quad / nu).astype(F32))
    return (-(logp.astype(F32)) * LOG2E_F32).astype(F32)


def studentt_nll_f64(x, sigma, nu):
    """Same formula evaluated in float64 from the float32 inputs (truth)."""
    x = np.asarray(x, np.float64)
    sigma = np.clip(np.asarray(sigma, np.float64), SIGMA_MIN, SIGMA_MAX)
    nu = np.clip(np.asarray(nu, np.float64), NU_MIN, NU_MAX)
    logC = (sp.gammaln((nu + 1.0) / 2.0) - sp.gammaln(nu / 2.0)
            - 0.5 * np.log(nu * math.pi) - np.log(sigma))
    logp = logC - (nu + 1.0) / 2.0 * np.log1p((x / sigma) ** 2 / nu)
    return -logp * LOG2E


def studentt_nll_grads_f64(x, sigma, nu, g=1.0):
    """Analytic gradients of g*nll (SURVEY 8(a'), checked there against autograd).

    Returns (dx, dsigma, dnu) per element; the clamp masks of distributions.py:23-24
    zero dsigma / dnu outside the closed clamp interval (torch.clamp semantics).
    """
    x = np.asarray(x, np.float64)
    s_raw = np.asarray(sigma, np.float64)
    n_raw = np.asarray(nu, np.float64)
    sc = np.clip(s_raw, SIGMA_MIN, SIGMA_MAX)
    nc = np.clip(n_raw, NU_MIN, NU_MAX)
    ms = ((s_raw >= SIGMA_MIN) & (s_raw <= SIGMA_MAX)).astype(np.float64)
    mn = ((n_raw >= NU_MIN) & (n_raw <= NU_MAX)).astype(np.float64)
    q = x * x
    den = nc * sc * sc + q
    u = q / (sc * sc * nc)
    dx = LOG2E * (nc + 1.0) * x / den
    dsig = LOG2E * (1.0 / sc - (nc + 1.0) * q / (sc * den)) * ms
    dnu = -LOG2E * (0.5 * sp.digamma((nc + 1.0) / 2.0) - 0.5 * sp.digamma(nc / 2.0)
                    - 1.0 / (2.0 * nc) - 0.5 * np.log1p(u)
                    + 0.5 * (nc + 1.0) * u / (nc * (1.0 + u))) * mn
    g = np.asarray(g, np.float64)
    return g * dx, g * dsig, g * dnu


# --------------------------------------------------------------------------- discretised Student-t (L2)
def studentt_bin_prob_f64(x, sigma, nu, mu=0.0):
    """P = T_nu(x+1/2) - T_nu(x-1/2), location mu, scale sigma (north_star; intended at
    eval_selfcontained_entropy.py:56-59).  Evaluated with same-side survival functions so the
    tails keep relative accuracy.  sigma/nu are clamped like the density (distributions.py:23-24)."""
    x = np.asarray(x, np.float64) - np.asarray(mu, np.float64)
    sc = np.clip(np.asarray(sigma, np.float64), SIGMA_MIN, SIGMA_MAX)
    nc = np.clip(np.asarray(nu, np.float64), NU_MIN, NU_MAX)
    lo = (x - 0.5) / sc
    hi = (x + 0.5) / sc
    nc = np.broadcast_to(nc, lo.shape)
    # survival S(t) = stdtr(nu, -t)
    p_right = sp.stdtr(nc, -lo) - sp.stdtr(nc, -hi)      # accurate when lo >= 0
    p_left = sp.stdtr(nc, hi) - sp.stdtr(nc, lo)         # accurate when hi <= 0
    p_mid = 1.0 - sp.stdtr(nc, -hi) - sp.stdtr(nc, lo)   # straddles 0
    return np.where(lo >= 0, p_right, np.where(hi <= 0, p_left, p_mid))


def studentt_cdfdiff_nll_f64(x, sigma, nu, mu=0.0):
    return -np.log2(studentt_bin_prob_f64(x, sigma, nu, mu))


# --------------------------------------------------------------------------- Gaussian z prior (L3)
def gaussian_nll_f32(x, log_sigma):
    """distributions.py:39-46, log_sigma is the per-channel parameter [C]; x is [B,C,h,w]."""
    x = np.asarray(x, F32)
    sigma = np.exp(np.asarray(log_sigma, F32)).astype(F32).reshape(1, -1, 1, 1)
    sigma = np.clip(sigma, F32(SIGMA_MIN), F32(SIGMA_MAX))
    var = (sigma ** 2).astype(F32)
    two_pi = F32(2 * math.pi)           # python double 2*pi -> fp32 scalar multiply
    logp = -F32(0.5) * np.log((two_pi * var).astype(F32)) - (F32(0.5) * (x ** 2).astype(F32)) / var
    return (-(logp.astype(F32)) * LOG2E_F32).astype(F32)


def gaussian_nll_f64(x, log_sigma):
    x = np.asarray(x, np.float64)
    sigma = np.clip(np.exp(np.asarray(log_sigma, np.float64)), SIGMA_MIN, SIGMA_MAX).reshape(1, -1, 1, 1)
    var = sigma ** 2
    return (0.5 * np.log(2 * math.pi * var) + 0.5 * x * x / var) * LOG2E


def gaussian_nll_grads_f64(x, log_sigma, g=1.0):
    """d(g*nll)/dx per element and d/dlog_sigma summed per channel (SURVEY 8(a') L3)."""
    x = np.asarray(x, np.float64)
    ls = np.asarray(log_sigma, np.float64)
    s_raw = np.exp(ls)
    sc = np.clip(s_raw, SIGMA_MIN, SIGMA_MAX)
    m = ((s_raw >= SIGMA_MIN) & (s_raw <= SIGMA_MAX)).astype(np.float64)
    var = (sc ** 2).reshape(1, -1, 1, 1)
    g = np.broadcast_to(np.asarray(g, np.float64), x.shape)
    dx = g * LOG2E * x / var
    # d nll / d sigma = LOG2E (1/sigma - x^2/sigma^3);  d sigma / d log_sigma = sigma (inside the clamp)
    dls = (g * LOG2E * (1.0 - x * x / var)).sum(axis=(0, 2, 3)) * m
    return dx, dls


# --------------------------------------------------------------------------- rate (R1)
def rate_bpp(nll_y, nll_z, n_img, h_img, w_img):
    """model.py:75-79: (sum nll_y + sum nll_z)/(N*H*W) clamped at 0 (float64 accumulation)."""
    r = (np.asarray(nll_y, np.float64).sum() + np.asarray(nll_z, np.float64).sum()) / (n_img * h_img * w_img)
    return max(r, 0.0)


# --------------------------------------------------------------------------- GDN / IGDN (G1, G2)
def gdn_effective_params(beta_param, gamma_weight):
    """layers.py:20-21: beta = beta_param^2 - 2^-18 ; gamma = weight^2 - 2^-18 (fp32, two roundings each)."""
    b = np.asarray(beta_param, F32)
    w = np.asarray(gamma_weight, F32).reshape(-1)
    return ((b * b).astype(F32) - REPARAM_OFFSET).astype(F32), ((w * w).astype(F32) - REPARAM_OFFSET).astype(F32)


def gdn_diag_fwd_f32(x, beta_param, gamma_weight, inverse=False):
    """layers.py:19-27 with IEEE single rounding at every step (no FMA contraction):
    x2=rn(x*x); p=rn(gamma*x2); s=rn(beta+p); d=sqrt_rn(s); y=rn(x/d) or rn(x*d)."""
    x = np.asarray(x, F32)
    beta, gamma = gdn_effective_params(beta_param, gamma_weight)
    x2 = (x * x).astype(F32)
    p = (gamma.reshape(1, -1, 1, 1) * x2).astype(F32)
    s = (beta.reshape(1, -1, 1, 1) + p).astype(F32)
    with np.errstate(invalid="ignore"):
        d = np.sqrt(s).astype(F32)       # NaN when s<0, as the reference would
    return (x * d).astype(F32) if inverse else (x / d).astype(F32)


def gdn_diag_bwd_f64(x, g, beta_param, gamma_weight, inverse=False):
    """Gradients of layers.py:19-27 (SURVEY 8(a') G1/G2), float64.

    Returns dx [B,C,H,W], dbeta_param [C], dweight [C] (chain through the squared
    re-parameterisation of layers.py:20-21 included)."""
    x = np.asarray(x, np.float64)
    g = np.asarray(g, np.float64)
    bp = np.asarray(beta_param, np.float64)
    wp = np.asarray(gamma_weight, np.float64).reshape(-1)
    beta32, gamma32 = gdn_effective_params(beta_param, gamma_weight)
    beta = beta32.astype(np.float64).reshape(1, -1, 1, 1)
    gamma = gamma32.astype(np.float64).reshape(1, -1, 1, 1)
    s = beta + gamma * x * x
    d = np.sqrt(s)
    if inverse:
        dx = g * (s + gamma * x * x) / d
        h = 0.5 * g * x / d
    else:
        dx = g * beta / (d * s)
        h = -0.5 * g * x / (d * s)
    dbeta = h.sum(axis=(0, 2, 3))
    dgamma = (h * x * x).sum(axis=(0, 2, 3))
    return dx, dbeta * 2.0 * bp, dgamma * 2.0 * wp


def gdn_dense_fwd_f64(x, beta, gamma, inverse=False):
    """north_star variant G3: s_i = beta_i + sum_j gamma_ij x_j^2 (oracle: F.conv2d(x^2, gamma.view(C,C,1,1), beta))."""
    x = np.asarray(x, np.float64)
    s = np.einsum("ij,bjhw->bihw", np.asarray(gamma, np.float64), x * x) + np.asarray(beta, np.float64).reshape(1, -1, 1, 1)
    d = np.sqrt(s)
    return x * d if inverse else x / d


def gdn_dense_bwd_f64(x, g, beta, gamma, inverse=False):
    """SURVEY 8(a') G3. Returns dx, dbeta [C], dgamma [C,C] w.r.t. the effective (non re-parameterised) beta/gamma."""
    x = np.asarray(x, np.float64)
    g = np.asarray(g, np.float64)
    gamma = np.asarray(gamma, np.float64)
    x2 = x * x
    s = np.einsum("ij,bjhw->bihw", gamma, x2) + np.asarray(beta, np.float64).reshape(1, -1, 1, 1)
    d = np.sqrt(s)
    if inverse:
        h = 0.5 * g * x / d
        direct = g * d
    else:
        h = -0.5 * g * x / (d * s)
        direct = g / d
    dx = direct + 2.0 * x * np.einsum("ij,bihw->bjhw", gamma, h)
    return dx, h.sum(axis=(0, 2, 3)), np.einsum("bihw,bjhw->ij", h, x2)


# --------------------------------------------------------------------------- symbols (I1) and support (T2/T3)
def symbols_and_support(q, tail=10):
    """eval_selfcontained_entropy.py:39-40,48 (z) / :52-53,62 (y), floor/ceil defect repaired:
    per patch  min = floor(min q) - tail,  max = ceil(max q) + tail,  sym = int32(q) - min."""
    q = np.asarray(q, F32)
    B = q.shape[0]
    mins = np.empty(B, np.int32)
    maxs = np.empty(B, np.int32)
    sym = np.empty(q.shape, np.int32)
    for b in range(B):
        mins[b] = int(math.floor(float(q[b].min()))) - tail
        maxs[b] = int(math.ceil(float(q[b].max()))) + tail
        sym[b] = q[b].astype(np.int32) - mins[b]
    return sym, mins, maxs


def pmf_to_uint16_cdf_spec(pmf_f32):
    """eval_selfcontained_entropy.py:17-23 along axis 0.

    torch's CPU cumsum accumulates float32 inputs in float64 and rounds every
    output back to float32 (ATen acc_type<float,false>=double); that is the order
    fixed here and in the CUDA kernel: c_k = f32(sum_{j<=k} (double)pmf_j)."""
    pmf = np.asarray(pmf_f32, F32)
    cdf = np.cumsum(pmf.astype(np.float64), axis=0).astype(F32)
    cdf = np.concatenate([np.zeros((1,) + cdf.shape[1:], F32), cdf], axis=0)
    cdf[-1] = np.maximum(cdf[-1], F32(1.0))
    scaled = np.clip((cdf * F32(65535.0)).astype(F32), F32(0), F32(65535.0))
    return scaled.astype(np.uint16)          # C-style truncation, like numpy astype in the reference


# --------------------------------------------------------------------------- in-kernel noise (Q1, perf mode)
def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11), vectorised over counters [n,4] uint32 with one key (k0,k1)."""
    c = np.array(counter, dtype=np.uint64).reshape(-1, 4).copy()
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = np.stack([(hi1 ^ c[:, 1] ^ k0) & mask, lo1, (hi0 ^ c[:, 3] ^ k1) & mask, lo0], axis=1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint32)


def philox_uniform_noise(n_elems, seed, offset):
    """The U(-1/2,1/2) draw of kernel K1 in SIC_QUANT_NOISE_PHILOX mode: vector v = i//4 is the counter low half,
    `offset` the high half, `seed` the key; 23 bits -> odd multiples of 2^-24."""
    nv = (n_elems + 3) // 4
    v = np.arange(nv, dtype=np.uint64)
    ctr = np.stack([v & np.uint64(0xFFFFFFFF), v >> np.uint64(32),
                    np.full(nv, offset & 0xFFFFFFFF, np.uint64), np.full(nv, (offset >> 32) & 0xFFFFFFFF, np.uint64)], axis=1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(-1)[:n_elems]
    return (((r >> np.uint32(9)).astype(np.float32) + F32(0.5)) * F32(2.0 ** -23) - F32(0.5)).astype(F32)
