"""Per-kernel time breakdown of one training step (torch.profiler/CUPTI): ours vs the eager PyTorch port on the same GPU."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import domain_specific_image_compression_b200 as sic
from domain_specific_image_compression_b200.trainer import FlatTrainer
from oracle import torch_port as TP
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "both"
cfg = bench.CONFIGS[os.environ.get("CFG", "cfg2")]
dev = torch.device("cuda", 0)
B = cfg["batch"]
x = bench.synthetic_batch(B, 256, 256, 42, dev)
if os.environ.get("CUDNN_BENCHMARK"): torch.backends.cudnn.benchmark = True

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def prof(fn, tag):
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
        for _ in range(3): fn()
        torch.cuda.synchronize()
    print(f"==== {tag}: top kernels over 3 steps")
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=int(os.environ.get("ROWS", "45")), max_name_column_width=70))

if which in ("both", "ours"):
    torch.manual_seed(42)
    model = sic.CompressionModel(N=cfg["N"], M=cfg["M"], min_nu=2.0).to(dev).train()
    with torch.no_grad():
        model.g_a.g_a[14].weight.mul_(40.0); model.h_a.h_a[6].weight.mul_(40.0); model.h_s.mlp_nu[2].bias.add_(1.5)
    if os.environ.get("CHANNELS_LAST"):
        model = model.to(memory_format=torch.channels_last); x = x.contiguous(memory_format=torch.channels_last)
    from domain_specific_image_compression_b200 import layers as _L
    _L.FUSE_FIRST_LAYER = bool(os.environ.get("FUSE_FIRST", "1") != "0" and os.environ.get("CHANNELS_LAST"))
    _L.FAST_LAST_LAYER = bool(os.environ.get("FAST_LAST", "1") != "0" and os.environ.get("CHANNELS_LAST"))
    from domain_specific_image_compression_b200 import model as _M
    _M.OVERLAP_HYPER_BRANCH = bool(os.environ.get("OVERLAP", "1") != "0" and os.environ.get("CHANNELS_LAST"))
    print("fused first layer", _L.FUSE_FIRST_LAYER, "gemm last layer", _L.FAST_LAST_LAYER, "hyper branch on side stream", _M.OVERLAP_HYPER_BRANCH)
    tr = FlatTrainer(model)
    def closure():
        out = model(x, "noise"); return sic.rate_distortion_loss(out, x, 10000.0, "msssim")[0]
    step = lambda: tr.step(closure)
    print("ours ms/step", timeit(step))
    def fwd_only():
        with torch.no_grad(): model(x, "noise")
    print("ours fwd only ms", timeit(fwd_only))
    prof(step, "ours")

if which in ("both", "port"):
    sd = TP.init_state(cfg["N"], cfg["M"], seed=42, device=dev)
    for k, v in sd.items():
        if not k.endswith(".gamma"): v.requires_grad_(True)
    opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-4)
    stepp = lambda: TP.train_step(sd, opt, x, 10000.0, "msssim", TP.multi_scale_ssim)
    print("eager port ms/step", timeit(stepp))
    prof(stepp, "eager port (reference op chains on the GPU)")
