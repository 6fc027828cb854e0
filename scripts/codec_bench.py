"""BASELINE.json configs[2]: compress/decompress round trip on 512x512 synthetic patches (N=128, M=192).
Checks bit-exactness of the round trip and times compress()/decompress() with the GPU and the host coder.

    python scripts/codec_bench.py [B]                                              one GPU, B patches
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/codec_bench.py [B]
        N GPUs, B patches PER GPU (weak scaling): the global batch of N*B patches is sharded by patch index
        (codec_parallel: rank r codes patches r, r+N, ...; no data-path collective), timed as the max over ranks between two barriers;
        one merged round trip through the host gather checks that the shards reassemble to the single-process result.
Prints one JSON object (rank 0)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)                      # NCCL banners etc. go to stderr; the JSON goes to the real stdout
import torch
import torch.distributed as dist
import domain_specific_image_compression_b200 as sic
from domain_specific_image_compression_b200 import codec_parallel as CP
import bench

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16
torch.manual_seed(42)
m = sic.CompressionModel(N=128, M=192, min_nu=2.0).to(dev).eval()
with torch.no_grad():
    m.g_a.g_a[14].weight.mul_(40.0); m.h_a.h_a[6].weight.mul_(40.0); m.h_s.mlp_nu[2].bias.add_(1.5)
NCHW = "--nchw" in sys.argv
if not NCHW:      # channels_last model and patches (as the training bench): cuDNN then runs without its NCHW<->NHWC conversion kernels,
    m = m.to(memory_format=torch.channels_last)      # 2.3 of 10.3 ms of GPU time per compress() call in the NCHW profile (r02ag)
x_all = bench.synthetic_batch(B * world, 512, 512, 7, dev)          # every rank holds the global batch; it codes its own patches
if not NCHW:
    x_all = x_all.contiguous(memory_format=torch.channels_last)
idx = CP.patch_indices(B * world, rank, world)
x = x_all[idx] if NCHW else x_all[idx].contiguous(memory_format=torch.channels_last)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, reps=5):
    fn(); barrier()
    ts = []
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter(); r = fn(); barrier(); dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t)
        ts.append(dt)
    ts.sort()
    return ts[len(ts) // 2], r


res = {"activation_layout": "NCHW" if "--nchw" in sys.argv else "channels_last", "n_gpus": world, "patches_per_gpu": B, "image": "512x512", "model": "N=128 M=192", "scaling": "weak",
       "sharding": "patch index modulo world size, no data-path collective (codec_parallel)"}
with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
    ref = m(x, "round")
for coder in ("gpu", "host"):
    tc, comp = timed(lambda: m.compress(x, coder=coder))
    td, xh = timed(lambda: m.decompress(comp, coder=coder))
    assert torch.equal(xh, ref["x_hat"].clamp(0, 1)), "round trip differs from forward()"
    nbytes = sum(len(s) for p in comp["strings"] for s in p)
    res[coder] = {"compress_ms": tc * 1e3, "decompress_ms": td * 1e3, "compress_patches_per_s": B * world / tc,
                  "decompress_patches_per_s": B * world / td, "bpp_real_rank0": nbytes * 8 / (B * 512 * 512)}
if world > 1:
    # the sharded public API end to end, including the host gather of the byte strings / reconstructions
    tcs, merged = timed(lambda: CP.compress_sharded(m, x_all), reps=3)
    tds, xh_all = timed(lambda: CP.decompress_sharded(m, merged), reps=3)
    assert len(merged["strings"]) == B * world
    assert torch.equal(xh_all[idx].to(dev), ref["x_hat"].clamp(0, 1)), "sharded round trip differs from forward() on this rank's patches"
    res["sharded_api_with_host_gather"] = {"compress_ms": tcs * 1e3, "decompress_ms": tds * 1e3,
                                           "compress_patches_per_s": B * world / tcs, "decompress_patches_per_s": B * world / tds,
                                           "note": "decompress: every rank ends with ALL reconstructions as a HOST tensor (NCCL all_gather on the "
                                                   "device, then one pageable D2H copy of the whole batch)"}
    tdd, xh_dev = timed(lambda: CP.decompress_sharded(m, merged, device_result=True), reps=3)
    assert torch.equal(xh_dev[idx], ref["x_hat"].clamp(0, 1))
    res["sharded_api_device_gather"] = {"decompress_ms": tdd * 1e3, "decompress_patches_per_s": B * world / tdd,
                                        "note": "every rank ends with all reconstructions on its device (one all_gather over NVLink)"}
with torch.no_grad():
    loss, R, D = sic.rate_distortion_loss(ref, x, 1.0, "mse")
res["bpp_estimated_density_rank0"] = float(R)
t_fwd, _ = timed(lambda: m(x, "round"))
res["forward_only_ms"] = t_fwd * 1e3
if rank == 0:
    print(json.dumps(res, indent=1), file=_OUT, flush=True)
if world > 1:
    barrier()
    dist.destroy_process_group()
