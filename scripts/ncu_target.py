"""Minimal launcher for ncu: each hot kernel at its roofline size (what kernel_bench.py quotes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from domain_specific_image_compression_b200 import functional as F
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"      # one target, or several separated by commas
which_set = set(which.split(","))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if which_set & {"all", "k1"}:
    y = torch.randn(16, 320, 128, 128, device=dev) * 3
    sg = torch.exp(torch.randn(16, 320, 1, 1, device=dev)); nu = torch.exp(torch.randn(16, 320, 1, 1, device=dev) + 1.5)
    for _ in range(reps):
        F.bottleneck(y, sg, nu, quant="noise")
        F.bottleneck(y, sg, nu, quant="round")
    yr = y.clone().requires_grad_(True); sr = sg.clone().requires_grad_(True); nr = nu.clone().requires_grad_(True)
    yt, nll, bits = F.bottleneck(yr, sr, nr, quant="noise")
    for _ in range(reps):
        torch.autograd.grad((bits, yt), (yr, sr, nr), (torch.ones_like(bits), torch.ones_like(yt)), retain_graph=True)
    if os.environ.get("CDF"):
        for _ in range(reps):
            F.bottleneck(y, sg, nu, quant="noise", lik="cdf_diff")
    del y, yr, yt, nll
if which_set & {"all", "gdn"}:
    for fmt in (torch.contiguous_format, torch.channels_last):
        x = torch.randn(16, 128, 256, 256, device=dev).contiguous(memory_format=fmt)
        g = torch.randn(16, 128, 256, 256, device=dev).contiguous(memory_format=fmt)
        beta = torch.sqrt(torch.rand(128, device=dev) + 0.5).requires_grad_(True)
        w = torch.sqrt(torch.rand(128, 1, 1, 1, device=dev) * 0.3 + 0.01).requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        for _ in range(reps):
            yv = F.gdn(xr, beta, w, False)
            torch.autograd.grad(yv, (xr, beta, w), g)
        del x, g, xr, yv
if which_set & {"all", "dense"}:
    x = torch.randn(16, 128, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
    beta = torch.sqrt(torch.rand(128, device=dev) + 0.5)
    gm = torch.sqrt(torch.rand(128, 128, device=dev) * 0.02 + torch.eye(128, device=dev) * 0.1 + 2.0 ** -18)
    for _ in range(reps):
        F.gdn_dense(x, beta, gm, False)
if which_set & {"dense192"}:
    x = torch.randn(8, 192, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
    beta = torch.sqrt(torch.rand(192, device=dev) + 0.5)
    gm = torch.sqrt(torch.rand(192, 192, device=dev) * 0.02 + torch.eye(192, device=dev) * 0.1 + 2.0 ** -18)
    for _ in range(reps):
        F.gdn_dense(x, beta, gm, False)
if which_set & {"cdf"}:
    y = torch.randn(16, 320, 128, 128, device=dev) * 3
    sg = torch.exp(torch.randn(16, 320, 1, 1, device=dev)); nu = torch.exp(torch.randn(16, 320, 1, 1, device=dev) + 1.5)
    for _ in range(reps):
        F.bottleneck(y, sg, nu, quant="noise", lik="cdf_diff")
if which_set & {"dense_bwd"}:
    for C, B in ((128, 16), (192, 8)):
        x = torch.randn(B, C, 256, 256, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        beta = torch.sqrt(torch.rand(C, device=dev) + 0.5).requires_grad_(True)
        gm = torch.sqrt(torch.rand(C, C, device=dev) * 0.02 + torch.eye(C, device=dev) * 0.1 + 2.0 ** -18).requires_grad_(True)
        for _ in range(reps):
            yv = F.gdn_dense(x, beta, gm, False)
            torch.autograd.grad(yv, (x, beta, gm), torch.randn_like(yv))
        del x, yv
if which_set & {"codec"}:
    import domain_specific_image_compression_b200 as sic
    torch.manual_seed(0)
    m = sic.CompressionModel(N=128, M=192, min_nu=2.0).to(dev).eval()
    with torch.no_grad():
        m.g_a.g_a[14].weight.mul_(40.0); m.h_a.h_a[6].weight.mul_(40.0); m.h_s.mlp_nu[2].bias.add_(1.5)
    xi = torch.rand(16, 3, 512, 512, device=dev)
    for _ in range(reps):
        comp = m.compress(xi)
        m.decompress(comp)
if which_set & {"ssim"}:
    from domain_specific_image_compression_b200 import losses
    a = torch.rand(16, 3, 256, 256, device=dev, requires_grad=True); b = torch.rand(16, 3, 256, 256, device=dev)
    for _ in range(reps):
        losses.multi_scale_ssim(a, b, 1.0, torch.tensor([0.3, 0.5, 0.2], device=dev)).backward()
torch.cuda.synchronize()
print("ok")
if which_set & {"conv0"}:
    x = torch.rand(16, 3, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(128, 3, 3, 3, device=dev) * 0.3).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(128, device=dev) * 0.2
    beta = torch.sqrt(torch.rand(128, device=dev) + 0.5); gw = torch.sqrt(torch.rand(128, 1, 1, 1, device=dev) * 0.3 + 0.01)
    ps = [t.clone().requires_grad_(True) for t in (w, bias, beta, gw)]
    for _ in range(reps):
        yv = F.conv0_gdn(x, *ps)
        torch.autograd.grad(yv, ps, torch.randn_like(yv))
    del x, yv
if which_set & {"deconv"}:
    a = torch.randn(16, 128, 128, 128, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(128, 3, 5, 5, device=dev) * 0.1).requires_grad_(True)
    bias = torch.randn(3, device=dev).requires_grad_(True)
    for _ in range(reps):
        yv = F.deconv_rgb(a, w, bias)
        torch.autograd.grad(yv, (a, w, bias), torch.randn_like(yv))
    del a, yv
if which_set & {"gdn128"}:
    x = torch.randn(16, 128, 128, 128, device=dev).contiguous(memory_format=torch.channels_last)
    g = torch.randn_like(x)
    beta = torch.sqrt(torch.rand(128, device=dev) + 0.5).requires_grad_(True)
    w = torch.sqrt(torch.rand(128, 1, 1, 1, device=dev) * 0.3 + 0.01).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    for _ in range(reps):
        for inv in (False, True):
            yv = F.gdn(xr, beta, w, inv)
            torch.autograd.grad(yv, (xr, beta, w), g)
    del x, g, xr, yv
if which_set & {"ssim"}:
    import domain_specific_image_compression_b200 as sic
    a = torch.rand(16, 3, 256, 256, device=dev).requires_grad_(True)
    b = torch.rand(16, 3, 256, 256, device=dev)
    for _ in range(reps):
        v = sic.losses.multi_scale_ssim(a, b, data_range=1.0, scale_weights=[0.3, 0.5, 0.2])
        torch.autograd.grad(v, a)
if which_set & {"codec"}:
    import domain_specific_image_compression_b200 as sic
    import bench
    torch.manual_seed(42)
    m = sic.CompressionModel(N=128, M=192, min_nu=2.0).to(dev).eval()
    with torch.no_grad():
        m.g_a.g_a[14].weight.mul_(40.0); m.h_a.h_a[6].weight.mul_(40.0); m.h_s.mlp_nu[2].bias.add_(1.5)
    xb = bench.synthetic_batch(4, 512, 512, 7, dev)
    for _ in range(reps):
        c = m.compress(xb)
        m.decompress(c)
if which_set & {"hyper"}:
    import domain_specific_image_compression_b200 as sic
    hs = sic.layers.HyperSynthesis(128, 192).to(dev)
    t = torch.relu(torch.randn(16, 128, 16, 16, device=dev)).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    for _ in range(reps):
        sg, nu = F.hyper_tail(t, hs.mlp_sigma, hs.mlp_nu, 2.0, 100.0)
        torch.autograd.grad((sg.sum() + nu.sum()), [t] + list(hs.mlp_sigma.parameters()) + list(hs.mlp_nu.parameters()))
