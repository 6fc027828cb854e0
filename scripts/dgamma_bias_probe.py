"""Does sic_gdn_dense_dgamma's single long-lived tensor-memory accumulator show the truncation bias the fused first layer's did
 at the full site size?  d(gamma) vs float64 at P = 65 536 ... 1 048 576 positions; prints max error and the mean signed
relative error (a bias shows as a consistent sign)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from domain_specific_image_compression_b200 import _lib
lib = _lib.load()
vp = lambda t: ctypes.c_void_p(t.data_ptr())
for C, P in ((128, 65536), (128, 262144), (128, 1048576), (192, 524288)):
    gen = torch.Generator(device="cuda").manual_seed(C + P)
    x = torch.randn(P, C, device="cuda", generator=gen) * 1.5
    h = torch.rand(P, C, device="cuda", generator=gen) * torch.rand(1, C, device="cuda", generator=gen)     # same-sign terms: worst case for a bias
    out = torch.empty(C, C, device="cuda")
    ws = torch.empty(lib.sic_gdn_dense_dgamma_workspace_bytes(P, C), dtype=torch.uint8, device="cuda")
    rc = lib.sic_gdn_dense_dgamma(vp(x), vp(h), P, C, vp(out), vp(ws), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.sic_last_error()
    ref = torch.zeros(C, C, dtype=torch.float64, device="cuda")
    for lo in range(0, P, 131072):
        ref += h[lo:lo + 131072].double().t() @ (x[lo:lo + 131072].double() ** 2)
    rel = (out.double() - ref) / ref
    print(f"C={C} P={P}: max |rel err| {float(rel.abs().max()):.3e}   mean signed rel err {float(rel.mean()):+.3e}", flush=True)
