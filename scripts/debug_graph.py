import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import domain_specific_image_compression_b200 as sic
from domain_specific_image_compression_b200 import functional as F
from domain_specific_image_compression_b200.trainer import FlatTrainer
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = bool(os.environ.get("BM"))

def try_capture(name, fn, warm=3):
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warm): fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        g.replay(); torch.cuda.synchronize()
        print("CAPTURE OK  :", name, flush=True)
    except Exception as e:
        print("CAPTURE FAIL:", name, type(e).__name__, str(e)[:200].replace("\n", " "), flush=True)
        traceback.print_exc(limit=8)
        torch.cuda.synchronize()

x = torch.randn(2, 16, 32, 32, device=dev, requires_grad=True)
b = torch.ones(16, device=dev, requires_grad=True); w = torch.full((16, 1, 1, 1), 0.3, device=dev, requires_grad=True)
try_capture("gdn fwd", lambda: F.gdn(x, b, w))
def gdn_fb():
    x.grad = None; b.grad = None; w.grad = None
    F.gdn(x, b, w).sum().backward()
try_capture("gdn fwd+bwd", gdn_fb)
y = torch.randn(2, 16, 8, 8, device=dev, requires_grad=True)
sg = torch.ones(2, 16, 1, 1, device=dev, requires_grad=True); nu = torch.full((2, 16, 1, 1), 5.0, device=dev, requires_grad=True)
try_capture("k1 fwd round", lambda: F.bottleneck(y, sg, nu, quant="round"))
try_capture("k1 fwd noise philox", lambda: F.bottleneck(y, sg, nu, quant="noise"))
def k1_fb():
    y.grad = None; sg.grad = None; nu.grad = None
    yt, nll, bits = F.bottleneck(y, sg, nu, quant="noise"); (bits.sum() + yt.sum()).backward()
try_capture("k1 fwd+bwd", k1_fb)
model = sic.CompressionModel(N=16, M=24, min_nu=2.0).to(dev).to(memory_format=torch.channels_last).train()
xb = torch.rand(2, 3, 64, 64, device=dev).contiguous(memory_format=torch.channels_last)
def fwd_loss_mse():
    for p in model.parameters(): p.grad = None
    out = model(xb, "noise"); l = sic.rate_distortion_loss(out, xb, 100.0, "mse")[0]; l.backward(); return l
try_capture("model fwd+loss(mse)+bwd", fwd_loss_mse)
def fwd_loss_ms():
    for p in model.parameters(): p.grad = None
    out = model(xb, "noise"); l = sic.rate_distortion_loss(out, xb, 100.0, "msssim")[0]; l.backward(); return l
try_capture("model fwd+loss(msssim)+bwd", fwd_loss_ms)
tr = FlatTrainer(model)
try_capture("trainer.step", lambda: tr.step(lambda: sic.rate_distortion_loss(model(xb, "noise"), xb, 100.0, "msssim")[0]))
