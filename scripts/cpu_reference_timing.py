"""Times the UNMODIFIED reference (imported from /root/reference, build container only) on the host cores:
SURVEY 8(d) items (i) cfg1 forward + loss, (ii) full step, (iii) isolated quantize+StudentT NLL and one GDN site.
    python scripts/cpu_reference_timing.py > profiles/r01_cpu_reference_container.json"""
import json, os, statistics, sys, time, types
REF = "/root/reference/code/modelv2"
sys.path.insert(0, REF)
sys.modules.setdefault("piq", types.ModuleType("piq"))
import torch
import distributions, layers, model as refmodel

cores = os.cpu_count()
torch.set_num_threads(cores)

def med(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return statistics.median(ts)

out = {"host": "build container", "cores": cores, "torch": torch.__version__, "kind": "reference (imported unmodified, dist='mse' because piq is absent)"}
torch.manual_seed(42)
m = refmodel.CompressionModel(N=128, M=192, spatial_params=False, min_nu=2, max_nu=100.0)
x = torch.rand(8, 3, 256, 256)
def fwd_loss():
    with torch.no_grad():
        o = m(x, "noise"); refmodel.rate_distortion_loss(o, x, 10000.0, "mse")
t = med(fwd_loss, 3); out["cfg1_forward_loss_B8"] = {"s": t, "patches_per_s": 8 / t}
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
def step():
    opt.zero_grad(set_to_none=True)
    o = m(x, "noise"); loss, _, _ = refmodel.rate_distortion_loss(o, x, 10000.0, "mse"); loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0); opt.step()
t = med(step, 3); out["cfg1_full_step_B8"] = {"s": t, "patches_per_s": 8 / t}
st = distributions.StudentT()
for shape in [(16, 192, 16, 16), (16, 320, 32, 32), (1, 320, 128, 128)]:
    y = torch.randn(*shape) * 3
    sg = torch.exp(torch.randn(shape[0], shape[1], 1, 1)).expand(*shape); nu = torch.exp(torch.randn(shape[0], shape[1], 1, 1) + 1.5).expand(*shape)
    def f():
        with torch.no_grad():
            yt = refmodel.CompressionModel.quantize(y, "noise"); st.neg_log2_prob(yt, sg, nu).sum()
    t = med(f); out[f"quantize_nll_sum_{'x'.join(map(str, shape))}"] = {"s": t, "Melem_per_s": y.numel() / t / 1e6}
g = layers.GDN(128)
xs = torch.randn(8, 128, 256, 256)
def gf():
    with torch.no_grad(): g(xs)
t = med(gf, 3); out["gdn_fwd_8x128x256x256"] = {"s": t, "GBps_algorithmic": 8 * xs.numel() / t / 1e9}
print(json.dumps(out, indent=1))
