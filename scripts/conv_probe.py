"""Which memory format does cuDNN serve best for each convolution of the model?  Times forward and backward (dgrad + wgrad) of every
conv / transposed-conv shape of the cfg2 step in NCHW and in channels_last, cudnn.benchmark on, TF32 allowed (the step's settings).
The round-1 launch list (profiles/r01s2_ncu_launches_bench_step.csv) shows the tiny hyper-network convolutions (16x16 ... 4x4
spatial) taking 130-210 us each in channels_last on cuDNN's legacy `convolve_common_engine_float_NHWC`.
    python scripts/conv_probe.py [--json out.json]"""
import json, sys, os
import torch, torch.nn as nn
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
B, N, M = 16, 128, 192
layers = [  # name, module factory, input shape
    ("g_a.0  conv 3->N 3x3 @256", lambda: nn.Conv2d(3, N, 3, 1, 1), (B, 3, 256, 256)),
    ("g_a.2  conv N->N 5x5s2 @256", lambda: nn.Conv2d(N, N, 5, 2, 2), (B, N, 256, 256)),
    ("g_a.4  conv N->N 3x3 @128", lambda: nn.Conv2d(N, N, 3, 1, 1), (B, N, 128, 128)),
    ("g_a.14 conv N->M 5x5s2 @32", lambda: nn.Conv2d(N, M, 5, 2, 2), (B, N, 32, 32)),
    ("h_a.0  conv M->N 3x3 @16", lambda: nn.Conv2d(M, N, 3, 1, 1), (B, M, 16, 16)),
    ("h_a.2  conv N->N 3x3 @16", lambda: nn.Conv2d(N, N, 3, 1, 1), (B, N, 16, 16)),
    ("h_a.4  conv N->N 5x5s2 @16", lambda: nn.Conv2d(N, N, 5, 2, 2), (B, N, 16, 16)),
    ("h_a.6  conv N->N 5x5s2 @8", lambda: nn.Conv2d(N, N, 5, 2, 2), (B, N, 8, 8)),
    ("h_s.0  deconv N->N @4", lambda: nn.ConvTranspose2d(N, N, 5, 2, 2, output_padding=1), (B, N, 4, 4)),
    ("h_s.2  deconv N->N @8", lambda: nn.ConvTranspose2d(N, N, 5, 2, 2, output_padding=1), (B, N, 8, 8)),
    ("g_s.0  deconv M->N @16", lambda: nn.ConvTranspose2d(M, N, 5, 2, 2, output_padding=1), (B, M, 16, 16)),
    ("g_s.2  conv N->N 3x3 @32", lambda: nn.Conv2d(N, N, 3, 1, 1), (B, N, 32, 32)),
    ("g_s.8  deconv N->N @64", lambda: nn.ConvTranspose2d(N, N, 5, 2, 2, output_padding=1), (B, N, 64, 64)),
    ("g_s.12 deconv N->3 @128", lambda: nn.ConvTranspose2d(N, 3, 5, 2, 2, output_padding=1), (B, N, 128, 128)),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def time_it(fn, reps=7):
    for _ in range(4): fn()
    ts = []
    for _ in range(reps):
        flush.zero_(); torch.cuda._sleep(400_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
rows = []
for name, make, shape in layers:
    row = {"layer": name}
    for fmt, tag in ((torch.contiguous_format, "nchw"), (torch.channels_last, "nhwc")):
        torch.manual_seed(0)
        m = make().to(dev).to(memory_format=fmt)
        x = torch.randn(*shape, device=dev).contiguous(memory_format=fmt).requires_grad_(True)
        y = m(x)
        g = torch.randn_like(y)
        row[f"fwd_{tag}_us"] = time_it(lambda: m(x))
        row[f"bwd_{tag}_us"] = time_it(lambda: torch.autograd.grad(y, (x, m.weight), g, retain_graph=True))
    rows.append(row)
    print(f"{name:30s} fwd nchw {row['fwd_nchw_us']:8.1f}  nhwc {row['fwd_nhwc_us']:8.1f}   bwd nchw {row['bwd_nchw_us']:8.1f}  nhwc {row['bwd_nhwc_us']:8.1f}", flush=True)
if "--json" in sys.argv:
    json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
