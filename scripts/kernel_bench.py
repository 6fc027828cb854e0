"""Kernel microbenchmark (BASELINE.json configs[4]): Student-t likelihood + rate kernel sweep, latent sizes 16x16x192 ...
128x128x320, plus the GDN sites of the training step; CUDA-event timing with an L2 flush between launches.

    python scripts/kernel_bench.py [--quick] [--only k1|gdn] [--json out.json]
Bytes are the ALGORITHMIC bytes of SURVEY.md 8(d) / BASELINE.md 3."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from domain_specific_image_compression_b200 import functional as F

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--only", default="")
ap.add_argument("--json", default="")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--eager-timing", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    PEAK = 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def time_it(fn, reps=args.reps, nbytes=0):
    """Kernels that stream >= 256 MB (2 x the 126 MB L2, walked front to back) are timed without a flush: "inputs larger than L2"
    (B200_PROFILING.md).  Flushing by writing 256 MB leaves 126 MB of dirty lines whose write-back is billed to the kernel.  Those
    launches are queued `reps` at a time behind a ~1 ms device-side spin (the host is then a full queue ahead: launch gaps are the
    device's own, as inside the step's CUDA graph); one event pair per batch, time / reps, median of 3.  --eager-timing keeps
    per-launch event pairs."""
    for _ in range(3): fn()
    big = nbytes >= (256 << 20)
    if big and not args.eager_timing:
        ts = []
        for _ in range(3):
            torch.cuda._sleep(2_000_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
        ts.sort()
        return ts[1] * 1e-3
    ts = []
    for _ in range(reps):
        if not big: flush.zero_()
        # a queued spin (~0.2 ms) lets the host run ahead of the device: without it the GPU sits idle between e0 and the kernel
        # while Python is still dispatching (~45 us for an autograd backward) and the idle time is billed to the kernel (r01:
        # gdn_bwd nhwc 295 us by events vs 252 us under ncu; every backward below 43 us "floor")
        torch.cuda._sleep(400_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3

rows = []
def report(name, shape, nbytes, t):
    r = {"kernel": name, "shape": list(shape), "elems": int(torch.Size(shape).numel()), "alg_bytes": nbytes, "us": t * 1e6, "gbs": nbytes / t / 1e9,
         "frac_of_measured_hbm_peak": nbytes / t / 1e9 / PEAK}
    rows.append(r)
    print(f"{name:34s} {str(tuple(shape)):24s} {t*1e6:9.1f} us {r['gbs']:8.1f} GB/s  {100*r['frac_of_measured_hbm_peak']:5.1f}% of {PEAK:.0f}", flush=True)

if args.only in ("", "k1"):
    sweep = [(C, h, w) for (C, h, w) in [(192, 16, 16), (192, 32, 32), (320, 32, 32), (192, 64, 64), (320, 64, 64), (320, 128, 128)]]
    batches = [16] if args.quick else [1, 16, 64]
    for B in batches:
        for (C, h, w) in sweep:
            n = B * C * h * w
            if n * 4 * 6 > 60e9: continue
            y = torch.randn(B, C, h, w, device=dev) * 3
            sg = torch.exp(torch.randn(B, C, 1, 1, device=dev)); nu = torch.exp(torch.randn(B, C, 1, 1, device=dev) + 1.5)
            report("k1_fwd density bcast noise", y.shape, 12 * n, time_it(lambda: F.bottleneck(y, sg, nu, quant="noise"), nbytes=12 * n))
            report("k1_fwd density bcast round", y.shape, 12 * n, time_it(lambda: F.bottleneck(y, sg, nu, quant="round"), nbytes=12 * n))
            yr = y.clone().requires_grad_(True); sr = sg.clone().requires_grad_(True); nr = nu.clone().requires_grad_(True)
            yt, nll, bits = F.bottleneck(yr, sr, nr, quant="noise")
            gb = torch.ones_like(bits); gy = torch.randn_like(yt)
            report("k1_bwd density bcast (bits+dy~)", y.shape, 12 * n, time_it(lambda: torch.autograd.grad((bits, yt), (yr, sr, nr), (gb, gy), retain_graph=True), nbytes=12 * n))
            if B == 16 or args.quick:
                ss = sg.expand_as(y).contiguous(); ns = nu.expand_as(y).contiguous()
                report("k1_fwd density spatial noise", y.shape, 20 * n, time_it(lambda: F.bottleneck(y, ss, ns, quant="noise"), nbytes=20 * n))
                try:
                    report("k1_fwd cdf_diff bcast noise", y.shape, 12 * n, time_it(lambda: F.bottleneck(y, sg, nu, quant="noise", lik="cdf_diff"), nbytes=12 * n))
                except Exception as e:
                    print("cdf_diff unavailable:", str(e)[:80])
                del ss, ns
            del y, yr, yt, nll, gy

if args.only in ("", "gdn"):
    for B, N in ([(16, 128)] if args.quick else [(16, 128), (64, 192)]):
        for hw in (256, 128, 64, 32):
            for fmt, tag in ((torch.contiguous_format, "nchw"), (torch.channels_last, "nhwc")):
                x = torch.randn(B, N, hw, hw, device=dev).contiguous(memory_format=fmt)
                g = torch.randn(B, N, hw, hw, device=dev).contiguous(memory_format=fmt)
                beta = torch.sqrt(torch.rand(N, device=dev) + 0.5).requires_grad_(True)
                w = torch.sqrt(torch.rand(N, 1, 1, 1, device=dev) * 0.3 + 0.01).requires_grad_(True)
                n = x.numel()
                for inv in (False, True):
                    nm = "igdn" if inv else "gdn"
                    report(f"{nm}_fwd {tag}", x.shape, 8 * n, time_it(lambda: F.gdn(x, beta, w, inv), nbytes=8 * n))
                    xr = x.clone().requires_grad_(True)
                    yv = F.gdn(xr, beta, w, inv)
                    report(f"{nm}_bwd {tag}", x.shape, 12 * n, time_it(lambda: torch.autograd.grad(yv, (xr, beta, w), g, retain_graph=True), nbytes=12 * n))
                    del xr, yv
                del x, g
if args.only in ("", "dense"):
    for B, N, hw in ([(16, 128, 256), (8, 192, 256)] if args.quick else [(16, 128, 256), (16, 128, 128), (16, 128, 64), (16, 64, 256), (8, 192, 256)]):
        x = torch.randn(B, N, hw, hw, device=dev).contiguous(memory_format=torch.channels_last)
        beta = torch.sqrt(torch.rand(N, device=dev) + 0.5)
        gm = torch.sqrt(torch.rand(N, N, device=dev) * 0.02 + torch.eye(N, device=dev) * 0.1 + 2.0 ** -18)
        n = x.numel()
        for variant, vname in ((0, "serial"), (1, "pipelined")):
            if N > 128 and variant == 0:
                continue
            for inv in (False, True):
                try:
                    t = time_it(lambda: F.gdn_dense(x, beta, gm, inv, variant), nbytes=8 * n)
                except Exception as e:   # keep the sweep going if one variant is broken on this box
                    print(f"dense {vname} inv={inv}: FAILED {e}", flush=True)
                    continue
                report(f"{'igdn' if inv else 'gdn'}_dense_fwd tcgen05 nhwc {vname}", x.shape, 8 * n, t)
                rows[-1]["tflops_tf32_2pass"] = 2 * 2 * N * N * (n // N) / t / 1e12
                print(f"    -> {rows[-1]['tflops_tf32_2pass']:.1f} TFLOP/s of tf32 MMA work (hi+lo passes)")
        # backward (12 B/elem algorithmic = read x, g; write dx): three tcgen05 launches (csrc/gdn_dense_bwd.cu passes
        # 1 and 2, csrc/gdn_dense_dgamma.cu) that move 40 B/elem
        try:
            xr = x.clone().requires_grad_(True)
            br, gr = beta.clone().requires_grad_(True), gm.clone().requires_grad_(True)
            yv = F.gdn_dense(xr, br, gr, False)
            go = torch.randn_like(yv)
            t = time_it(lambda: torch.autograd.grad(yv, (xr, br, gr), go, retain_graph=True), reps=5, nbytes=12 * n)
            report("gdn_dense_bwd tcgen05 3-pass", x.shape, 12 * n, t)
            t = time_it(lambda: torch.autograd.grad(yv, (xr, br), go, retain_graph=True), reps=5, nbytes=12 * n)
            report("gdn_dense_bwd tcgen05 (dx, dbeta only)", x.shape, 12 * n, t)
            del xr, yv, go
        except Exception as e:
            print(f"dense bwd: FAILED {e}", flush=True)
        del x
if args.only in ("", "conv0"):
    # N2: first analysis layer fused (conv 3->C 3x3 + bias + GDN, one kernel each way) vs the unfused cuDNN conv + GDN kernel.
    # Algorithmic bytes: forward = image in + y out; backward = image + grad_y in (parameter gradients out are KBs).
    import torch.nn.functional as TF
    torch.backends.cudnn.benchmark = True
    for (B, C, H, W) in [(16, 128, 256, 256), (8, 192, 256, 256)]:
        try:
            x = torch.rand(B, 3, H, W, device=dev).contiguous(memory_format=torch.channels_last)
            w = (torch.randn(C, 3, 3, 3, device=dev) * 0.3).contiguous(memory_format=torch.channels_last)
            bias = torch.randn(C, device=dev) * 0.2
            beta = torch.sqrt(torch.rand(C, device=dev) + 0.5); gw = torch.sqrt(torch.rand(C, 1, 1, 1, device=dev) * 0.3 + 0.01)
            n = B * C * H * W
            fb, bb = 4 * n + 4 * x.numel(), 4 * n + 4 * x.numel()
            report("conv0+gdn fused fwd (tcgen05)", (B, C, H, W), fb, time_it(lambda: F.conv0_gdn(x, w, bias, beta, gw), nbytes=fb))
            ps = [t.clone().requires_grad_(True) for t in (w, bias, beta, gw)]
            y = F.conv0_gdn(x, *ps); go = torch.randn_like(y)
            report("conv0+gdn fused bwd (tcgen05 x2)", (B, C, H, W), bb, time_it(lambda: torch.autograd.grad(y, ps, go, retain_graph=True), nbytes=bb))
            del y
            unf = lambda: F.gdn(TF.conv2d(x, w, None, 1, 1), beta, gw, False, bias=bias)
            report("conv0 cuDNN + gdn kernel fwd", (B, C, H, W), fb, time_it(unf, nbytes=fb))
            ps = [t.clone().requires_grad_(True) for t in (w, bias, beta, gw)]
            y = F.gdn(TF.conv2d(x, ps[0], None, 1, 1), ps[2], ps[3], False, bias=ps[1])
            report("conv0 cuDNN + gdn kernel bwd", (B, C, H, W), bb, time_it(lambda: torch.autograd.grad(y, ps, go, retain_graph=True), nbytes=bb))
            del y, go, x
        except Exception as e:
            print(f"conv0 {B}x{C}: FAILED {type(e).__name__}: {e}", flush=True)
if args.only in ("", "deconv"):
    # N2, synthesis side: last layer deconv(N, 3) as library GEMM + gather kernels vs cuDNN's conv_transpose2d (legacy 3-band engines).
    # Algorithmic bytes: forward = a in + x_hat out; backward = a, grad in + d(a) out.
    import torch.nn.functional as TF
    torch.backends.cudnn.benchmark = True
    for (B, N, H, W) in [(16, 128, 128, 128), (8, 192, 128, 128)]:
        try:
            a = torch.randn(B, N, H, W, device=dev).contiguous(memory_format=torch.channels_last)
            w = torch.randn(N, 3, 5, 5, device=dev) * 0.1
            bias = torch.randn(3, device=dev)
            n, no = a.numel(), B * 3 * 4 * H * W
            fb, bb = 4 * (n + no), 4 * (2 * n + no)
            report("deconv_rgb GEMM+col2im fwd", (B, N, H, W), fb, time_it(lambda: F.deconv_rgb(a, w, bias), nbytes=1 << 30))
            ps = [t.clone().requires_grad_(True) for t in (a, w, bias)]
            y = F.deconv_rgb(*ps); go = torch.randn_like(y)
            report("deconv_rgb im2col+2 GEMM bwd", (B, N, H, W), bb, time_it(lambda: torch.autograd.grad(y, ps, go, retain_graph=True), nbytes=1 << 30))
            del y
            report("deconv cuDNN fwd", (B, N, H, W), fb, time_it(lambda: TF.conv_transpose2d(a, w, bias, 2, 2, 1), nbytes=1 << 30))
            ps = [t.clone().requires_grad_(True) for t in (a, w, bias)]
            y = TF.conv_transpose2d(*ps, 2, 2, 1)
            report("deconv cuDNN bwd", (B, N, H, W), bb, time_it(lambda: torch.autograd.grad(y, ps, go, retain_graph=True), nbytes=1 << 30))
            del y, go, a
        except Exception as e:
            print(f"deconv {B}x{N}: FAILED {type(e).__name__}: {e}", flush=True)
if args.json:
    json.dump({"peak_gbs": PEAK, "rows": rows}, open(args.json, "w"), indent=1)
