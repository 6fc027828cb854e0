"""BASELINE.json configs[2]: compress/decompress round trip on 512x512 synthetic patches (N=128, M=192).
Checks bit-exactness of the round trip and times compress()/decompress() with the GPU and the host coder."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import domain_specific_image_compression_b200 as sic
import bench

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(42)
m = sic.CompressionModel(N=128, M=192, min_nu=2.0).to(dev).eval()
with torch.no_grad():
    m.g_a.g_a[14].weight.mul_(40.0); m.h_a.h_a[6].weight.mul_(40.0); m.h_s.mlp_nu[2].bias.add_(1.5)
x = bench.synthetic_batch(B, 512, 512, 7, dev)

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], r

res = {"batch": B, "image": "512x512", "model": "N=128 M=192"}
with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
    ref = m(x, "round")
for coder in ("gpu", "host"):
    tc, comp = timed(lambda: m.compress(x, coder=coder))
    td, xh = timed(lambda: m.decompress(comp, coder=coder))
    assert torch.equal(xh, ref["x_hat"].clamp(0, 1)), "round trip differs from forward()"
    nbytes = sum(len(s) for p in comp["strings"] for s in p)
    res[coder] = {"compress_ms": tc * 1e3, "decompress_ms": td * 1e3, "compress_patches_per_s": B / tc, "decompress_patches_per_s": B / td,
                  "bpp_real": nbytes * 8 / (B * 512 * 512)}
with torch.no_grad():
    loss, R, D = sic.rate_distortion_loss(ref, x, 1.0, "mse")
res["bpp_estimated_density"] = float(R)
t_fwd, _ = timed(lambda: m(x, "round"))
res["forward_only_ms"] = t_fwd * 1e3
print(json.dumps(res, indent=1))
