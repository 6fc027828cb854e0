import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from domain_specific_image_compression_b200 import functional as F
from oracle import numpy_ref as R
rng = np.random.default_rng(21)
shape=(2,6,4,4); B,C=2,6
sig = np.exp(rng.uniform(np.log(2e-2), np.log(50.0), (B,C,1,1))).astype(np.float32)
nu = np.clip(np.exp(rng.uniform(np.log(2.0), np.log(100.0), (B,C,1,1))),2.05,95).astype(np.float32)
y = (rng.standard_t(3, shape) * sig * rng.choice([1.0, 4.0], shape)).astype(np.float32)
dev=lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
yd, sd, nd = dev(y).requires_grad_(True), dev(sig).requires_grad_(True), dev(nu).requires_grad_(True)
_, nll, _ = F.bottleneck(yd, sd, nd, quant="none", lik="cdf_diff")
nll.sum().backward()
f = lambda yy, ss, nn: R.studentt_cdfdiff_nll_f64(yy, ss, nn)
y64, s64, n64 = y.astype(np.float64), sig.astype(np.float64), nu.astype(np.float64)
e=1e-5
dx = (f(y64 + e, s64, n64) - f(y64 - e, s64, n64)) / (2 * e)
got=yd.grad.cpu().numpy()
err=np.abs(got-dx); idx=np.argsort(err.ravel())[::-1][:12]
S=np.broadcast_to(sig,shape).ravel(); N=np.broadcast_to(nu,shape).ravel()
for i in idx:
    print(f"y={y.ravel()[i]:.5g} sigma={S[i]:.4g} nu={N[i]:.4g} lo={(y.ravel()[i]-.5)/S[i]:.4g} hi={(y.ravel()[i]+.5)/S[i]:.4g} got={got.ravel()[i]:.6g} want={dx.ravel()[i]:.6g} nll={nll.detach().cpu().numpy().ravel()[i]:.5g} ref_nll={f(y64,s64,n64).ravel()[i]:.5g}")
