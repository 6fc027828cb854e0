"""CUPTI breakdown of compress() / decompress() on a 16 x 512^2 batch (cfg3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import domain_specific_image_compression_b200 as sic
import bench
dev = torch.device("cuda", 0)
torch.manual_seed(42)
m = sic.CompressionModel(N=128, M=192, min_nu=2.0).to(dev).eval()
with torch.no_grad():
    m.g_a.g_a[14].weight.mul_(40.0); m.h_a.h_a[6].weight.mul_(40.0); m.h_s.mlp_nu[2].bias.add_(1.5)
x = bench.synthetic_batch(16, 512, 512, 7, dev)
for _ in range(3):
    c = m.compress(x); m.decompress(c)
torch.cuda.synchronize()
for name, fn in (("compress", lambda: m.compress(x)), ("decompress", lambda: m.decompress(c))):
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
        for _ in range(3): fn()
        torch.cuda.synchronize()
    print(f"==== {name}: 3 calls")
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
