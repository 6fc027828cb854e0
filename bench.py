#!/usr/bin/env python3
"""bench.py — train-step patches/s of the modelv2 Student-t hyperprior autoencoder on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = the reference's training step (train.py:193-204 without AMP: forward 'noise' -> rate_distortion_loss(msssim)
-> backward -> clip 1.0 -> Adam) on one batch of synthetic 256x256 3-band patches; per-GPU batch fixed (weak scaling),
one flat-bucket NCCL all-reduce per step for N > 1.  Prints ONE JSON line (rank 0).
  value     patches/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public API with the batch coming from pinned HOST memory every step and the loss read
            back to the host every step (trainer.HostFedLoop: copies on a copy stream, loss handed back one step late)
  roofline  dominant own kernel of the step (GDN backward in the step's layout): algorithmic bytes / CUDA-event time,
            measured in this run; `traffic` from the committed ncu capture of the SAME register allocation
            (profiles/ncu_traffic.json vs sic_kernel_registers of the loaded library), else null
  cpu_baseline  the reference's eager op chains (oracle/torch_port.py, kind "port": /root/reference itself cannot travel to
            the GPU box; the port is pinned bit-for-bit against the imported reference by tests/test_oracle_golden.py) on
            the host cores, bounded sample, median step
  gpu_eager_baseline  the same eager op chains on THIS GPU (fp32, NCHW, same cuDNN flags): the bar the fused kernels must beat
  clocks    SM clock (median) and throttle reasons of the samples stamped INSIDE the timed region, from an in-process NVML thread
            (10 ms period) started before the warm-up; `value` and `e2e` are each timed as W warm-up steps + K steps from an idle
            device (--e2e-idle-s between them); --settle-ms gives the sustained, power-capped figure instead of the first 0.1 s
--impl reference: only the CPU arm (rank 0), same metric/config/unit, the caller's --steps/--warmup honoured.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_REAL_STDOUT = sys.stdout

CONFIGS = {
    # BASELINE.json configs[1]: full training step batch 16 fp32 on 1xB200 (N=128, M=192)
    "cfg2": dict(N=128, M=192, batch=16, H=256, W=256),
    # BASELINE.json configs[3]: larger model, batch 64 per GPU
    "cfg4": dict(N=192, M=320, batch=64, H=256, W=256),
}
LAMBDA_RD = 10000.0      # config.py:38
METRIC = "train_step_patches_per_sec"
UNIT = "patches/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--e2e-idle-s", type=float, default=1.0, help="idle time between the device-resident measurement and the end-to-end one "
                    "(both then start from an idle device, W warm-up steps, K timed steps)")
    ap.add_argument("--settle-ms", type=float, default=0.0, help="0 (default, the bench contract): exactly W warm-up steps, then K timed ones.  > 0: "
                    "keep running untimed steps for about this long first - the sustained figure under the board's power cap (r02aw: "
                    "SM clock 1755 MHz and 5.39 ms/step after 0.4 s of back-to-back steps, against 1875-1940 MHz and 5.14 ms in the first 0.15 s)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--dist", default="msssim", choices=["msssim", "mse"])
    ap.add_argument("--ref-batch", type=int, default=0, help="patches per CPU step of the CPU arms (bounded sample); 0 = 8 for "
                    "N<=128 (BASELINE.json configs[0] runs the CPU case at batch 8), 4 for the larger model")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--gdn", default="diag", choices=["diag", "dense"], help="diag: the reference's path (layers.py:21); dense: all 13 "
                    "GDN/IGDN sites use the C x C gamma on tcgen05 tensor cores (north_star's contraction), forward and backward")
    ap.add_argument("--bucket-mb", type=float, default=None, help="gradient bucket size for the overlapped all-reduce (default: 6 MB at N > 1, "
                    "one bucket at N = 1 where there is nothing to overlap)")
    ap.add_argument("--nchw", action="store_true", help="keep activations NCHW (default: torch.channels_last, which saves cuDNN's "
                    "internal NCHW<->NHWC transposes; GDN kernels run natively in either layout)")
    ap.add_argument("--no-cudnn-benchmark", action="store_true", help="cuDNN autotune is on by default (warm-up steps absorb it)")
    ap.add_argument("--cudnn-benchmark-limit", type=int, default=-1, help="torch.backends.cudnn.benchmark_limit (PyTorch default 10 "
                    "candidates per convolution; 0 = try every engine configuration cuDNN's heuristics return); -1 leaves it alone")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying one CUDA graph")
    ap.add_argument("--no-fuse-first-layer", action="store_true", help="keep g_a's first layer as cuDNN conv + GDN kernel (default: the fused "
                    "conv 3->N 3x3 + bias + GDN tcgen05 kernel, forward and backward: layers.FUSE_FIRST_LAYER)")
    ap.add_argument("--no-overlap-hyper", action="store_true", help="run the hyperprior branch in line with the synthesis transform "
                    "(default: on a side stream next to it, forward and backward: model.OVERLAP_HYPER_BRANCH)")
    ap.add_argument("--no-fast-last-layer", action="store_true", help="keep g_s's last layer deconv(N,3) on cuDNN (default: library GEMM + "
                    "the col2im / im2col gather kernels: layers.FAST_LAST_LAYER)")
    ap.add_argument("--pad-rgb", type=int, default=0, choices=[0, 4, 8], help="zero-pad the image-side channel axis of the first conv / "
                    "last transposed conv to 4 or 8 so cuDNN can use tensor-core kernels there (layers.PAD_RGB_CHANNELS; identity "
                    "in exact arithmetic, off by default)")
    return ap.parse_args()


def synthetic_batch(batch, H, W, seed, device):
    """seeded uniform noise, low-passed so the latents are not degenerate; values in [0,1]."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, H // 8, W // 8, generator=g)
    x = torch.nn.functional.interpolate(x, size=(H, W), mode="bilinear", align_corners=False).clamp(0, 1)
    return x.to(device)


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's eager op chains on the host cores
def ref_batch(args, cfg):
    return args.ref_batch if args.ref_batch > 0 else (8 if cfg["N"] <= 128 else 4)


def cpu_reference(cfg, steps, warmup, batch, dist_name):
    """-> (patches/s from the MEDIAN step, median ms/step, cores, all step times).  Median, not mean: the first steps after the
    warm-up still pay allocator / oneDNN primitive-cache effects, which made two runs of the same code differ by 1.5x in round 1."""
    import torch
    from oracle import torch_port as TP
    multi_scale_ssim = TP.multi_scale_ssim         # the oracle's own restatement of piq: no product code on the CPU arm
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = TP.init_state(cfg["N"], cfg["M"], seed=42, device="cpu")
    for k, v in sd.items():
        if not k.endswith(".gamma"):
            v.requires_grad_(True)
    opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-4)
    x = synthetic_batch(batch, cfg["H"], cfg["W"], 42, "cpu")
    for _ in range(warmup):
        TP.train_step(sd, opt, x, LAMBDA_RD, dist_name, multi_scale_ssim)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        TP.train_step(sd, opt, x, LAMBDA_RD, dist_name, multi_scale_ssim)
        times.append(time.perf_counter() - t0)
    med = sorted(times)[len(times) // 2]
    return batch / med, med * 1e3, cores, [t * 1e3 for t in times]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    steps, warmup, rb = max(1, args.steps), max(0, args.warmup), ref_batch(args, cfg)
    # bounded sample: rb patches per step (~1.2 s of CPU work per step on 16 cores), so the default 20 + 5 steps end in ~30 s
    value, ms, cores, times = cpu_reference(cfg, steps, warmup, rb, args.dist)
    sample = (f"{rb} patches/step x {steps} timed steps (after {warmup} warm-up steps) of the {args.config} training step; value = "
              f"patches / median step; oracle/torch_port.py eager fp32 (bit-equal to the imported reference on CPU)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, cfg), "batch_per_step": rb, "parallelism": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "step_ms_min_median_max": [min(times), ms, max(times)]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def gpu_eager_reference(cfg, dev, dist_name, batch, steps=3, warmup=2):
    """The reference's eager op chains (oracle/torch_port.py) on THIS GPU: fp32, NCHW (the reference's layout), TF32 convs allowed
    and cuDNN autotune on exactly like our arm, no AMP (BASELINE configs are fp32).  CUDA-event timed, median step."""
    import torch
    from oracle import torch_port as TP
    sd = TP.init_state(cfg["N"], cfg["M"], seed=42, device=dev)
    for k, v in sd.items():
        if not k.endswith(".gamma"):
            v.requires_grad_(True)
    opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-4)
    x = synthetic_batch(batch, cfg["H"], cfg["W"], 42, dev)
    for _ in range(warmup):
        TP.train_step(sd, opt, x, LAMBDA_RD, dist_name, TP.multi_scale_ssim)
    torch.cuda.synchronize()
    times = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        TP.train_step(sd, opt, x, LAMBDA_RD, dist_name, TP.multi_scale_ssim)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    med = sorted(times)[len(times) // 2]
    del sd, opt, x
    torch.cuda.empty_cache()
    return {"value": batch / (med * 1e-3), "unit": UNIT, "ms_per_step": med, "batch": batch, "steps": steps, "warmup": warmup,
            "what": "oracle/torch_port.py (the reference's op chains, eager PyTorch) on this GPU: fp32, NCHW, cudnn.benchmark on, TF32 convs allowed, "
                    "no CUDA graph, restated MS-SSIM; inputs resident"}


def workload_name(args, cfg):
    return (f"{args.config}: modelv2 Student-t hyperprior AE N={cfg['N']} M={cfg['M']}, full training step "
            f"(fwd 'noise' + {args.dist} RD loss + bwd + clip + Adam), batch {cfg['batch']}/GPU, fp32, {cfg['H']}x{cfg['W']} synthetic 3-band patches")


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).

    In-process NVML (pynvml), polled every 10 ms by a thread that is started BEFORE the warm-up steps: the round's earlier version
    spawned `nvidia-smi -lms 100` at the start of the timed region, and that process's NVML initialisation (it enumerates every GPU of
    the box) stalled work submission for a few ms of a 100 ms region - the device-resident loop measured 3-6 % slower than the
    end-to-end loop timed right after it (r02ay: 5.14 vs 4.89 ms, r02ba: 5.25 vs 4.95 ms).  Only samples stamped inside
    [mark_begin, mark_end] are reported.  Falls back to the nvidia-smi process (also started early) when pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.handle = index, [], None, None, None
        self.t0 = self.t1 = None
        self._stop = threading.Event()
        self.thread = None

    def _open_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        handle = None
        try:                                                     # the CUDA device may be remapped (CUDA_VISIBLE_DEVICES): go by UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.nvml, self.handle = pynvml, handle
        self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        get = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {}
        for name, a, b in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                           ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                           ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap")):
            bits[name] = getattr(pynvml, a, None) or getattr(pynvml, b)
        self._get_reasons, self._bits = get, bits

    def _poll_nvml(self):
        nv, h = self.nvml, self.handle
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(self._get_reasons(h))
                self.rows.append((time.perf_counter(), sm, self.max_sm, [n for n, b in self._bits.items() if mask & b]))
            except Exception:
                pass
            self._stop.wait(0.010)

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(r[0]), float(r[1]),
                                  [n for n, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def start(self):
        try:
            self._open_nvml()
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            self.source = "nvml in-process, 10 ms period"
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read_smi, daemon=True).start()
            self.source = "nvidia-smi -lms 50"
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.proc is not None:
            time.sleep(0.1)
            self.proc.terminate()
        rows = list(self.rows)
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        note = None
        if not inside and rows and self.t0 is not None:          # a region shorter than the sampling period: the nearest sample
            mid = 0.5 * (self.t0 + (self.t1 or self.t0))
            inside, note = [min(rows, key=lambda r: abs(r[0] - mid))], "no sample fell inside the timed region: nearest one reported"
        sm = sorted(r[1] for r in inside)
        reasons = sorted({n for r in inside for n in r[3]})
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in inside), default=None),
               "reasons": reasons, "samples": len(sm), "source": self.source}
        if note:
            out["note"] = note
        return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import domain_specific_image_compression_b200 as sic
    from domain_specific_image_compression_b200 import functional as F_sic
    from domain_specific_image_compression_b200.trainer import FlatTrainer, HostFedLoop

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the product has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"[bench] warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    cfg = CONFIGS[args.config]
    B, H, W = cfg["batch"], cfg["H"], cfg["W"]
    # kernel rooflines first: they draw from torch's CUDA generator, which must not be touched between graph replays
    fused_first = not (args.no_fuse_first_layer or args.pad_rgb or args.gdn == "dense" or args.nchw)
    roof, kernels = kernel_rooflines(cfg, dev, channels_last=not args.nchw, fused_first_layer=fused_first) if rank == 0 else (None, None)
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    if args.cudnn_benchmark_limit >= 0:
        torch.backends.cudnn.benchmark_limit = args.cudnn_benchmark_limit
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager_baseline:
        try:
            gpu_eager = gpu_eager_reference(cfg, dev, args.dist, min(B, 16))
        except Exception as e:                                   # a comparator must never take the measurement down with it
            gpu_eager = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
            torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    torch.manual_seed(42)                                      # config.py:32
    model = sic.CompressionModel(N=cfg["N"], M=cfg["M"], spatial_params=False, min_nu=2.0, max_nu=100.0).to(dev)
    with torch.no_grad():                                      # 'spread' init so latents are not all zero
        model.g_a.g_a[14].weight.mul_(40.0)
        model.h_a.h_a[6].weight.mul_(40.0)
        model.h_s.mlp_nu[2].bias.add_(1.5)
    if args.gdn == "dense":                                    # north_star's G3 at every site; gamma (C x C) becomes the live parameter
        from domain_specific_image_compression_b200.layers import GDN as _GDN
        for mod in model.modules():
            if isinstance(mod, _GDN):
                mod.dense = True
    model.train()
    fmt = torch.contiguous_format if args.nchw else torch.channels_last
    from domain_specific_image_compression_b200 import layers as _layers
    if args.pad_rgb:
        _layers.PAD_RGB_CHANNELS = args.pad_rgb
    _layers.FUSE_FIRST_LAYER = not (args.no_fuse_first_layer or args.pad_rgb or args.gdn == "dense" or args.nchw)
    _layers.FAST_LAST_LAYER = not (args.no_fast_last_layer or args.pad_rgb or args.nchw)
    from domain_specific_image_compression_b200 import model as _model
    _model.OVERLAP_HYPER_BRANCH = not args.no_overlap_hyper
    model = model.to(memory_format=fmt)
    trainer = FlatTrainer(model, lr=1e-4, betas=(0.9, 0.999), grad_clip=1.0, bucket_bytes=None if args.bucket_mb is None else int(args.bucket_mb * (1 << 20)))
    x_dev = synthetic_batch(B, H, W, 42 + rank, dev).contiguous(memory_format=fmt)
    x_host = x_dev.cpu().pin_memory()
    x_stage = torch.empty_like(x_dev)

    def closure_on(xin):
        def closure():
            out = model(xin, quant_mode="noise")
            loss, _, _ = sic.rate_distortion_loss(out, xin, lambda_rd=LAMBDA_RD, dist=args.dist)
            return loss
        return closure

    graph_note = "eager launches"
    run_static = None
    x_stage.copy_(x_dev)
    launches_per_step = None
    if not args.no_graph:
        try:
            n_cap = max(3, args.warmup)
            lc0 = F_sic.launch_count
            run_static = trainer.capture(closure_on(x_stage), warmup=n_cap)
            launches_per_step = (F_sic.launch_count - lc0) // (n_cap + 1)     # our kernels recorded per replayed step
            graph_note = "whole step (zero_grad..Adam, incl. the bucketed all-reduces overlapped with backward) replayed as one CUDA graph"
        except Exception as e:                                   # fall back loudly, never silently
            print(f"[bench] CUDA graph capture failed, running eagerly: {type(e).__name__}: {e}", file=sys.stderr)
            run_static = None
    x_stage.copy_(x_dev)

    def step_resident():                                         # inputs already resident in HBM (x_stage holds the batch)
        return run_static() if run_static is not None else trainer.step(closure_on(x_stage))

    # end to end: every step's batch comes from pinned host memory and every step's loss goes back to the host; the copies run on
    # a copy stream next to the previous / next step's kernels and the loss is read one step late (trainer.HostFedLoop)
    feed = HostFedLoop(step_resident, x_stage)
    e2e_losses = []

    def step_e2e():
        if not feed._staged:
            feed.stage(x_host)
        prev = feed.step(x_host)                                 # runs this step, starts the next batch's H2D, returns the previous loss
        if prev is not None:
            e2e_losses.append(prev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            allms = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            per_rank_ms.clear()
            per_rank_ms.extend(float(v) / steps for v in allms)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    per_rank_ms = []
    W_, K = max(3, args.warmup), max(1, args.steps)
    if os.environ.get("SIC_BENCH_START_AT"):                     # diagnosis only (scripts/gpu_duo.sh): independent processes on different
        torch.cuda.synchronize()                                 # GPUs of one box start their warm-up at the same wall-clock time
        time.sleep(max(0.0, float(os.environ["SIC_BENCH_START_AT"]) - time.time()))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                          # before the warm-up: NVML initialisation stays out of the timed region
    ms_warm = timed(step_resident, W_)                           # the W warm-up steps (timed only to size the settle phase below)
    # --settle-ms > 0: keep stepping, untimed, for about that long before the timed region (the sustained, power-capped figure); the
    # count comes from the max-over-ranks warm-up time, so every rank runs the same number of steps (each one holds collectives).
    # The timed region is exactly K steps either way.
    extra_warm = 0
    if args.settle_ms > 0:
        extra_warm = int(min(500, max(0, round(args.settle_ms / max(ms_warm / W_, 1e-3)))))
        for _ in range(extra_warm):
            step_resident()
    l0 = F_sic.launch_count
    torch.cuda.profiler.start()                              # ncu --profile-from-start off isolates the timed region (all threads)
    sampler.mark_begin()
    ms_total = timed(step_resident, K)
    sampler.mark_end()
    torch.cuda.profiler.stop()
    launches = F_sic.launch_count - l0 if run_static is None else launches_per_step * K
    clocks = sampler.stop() if rank == 0 else None
    # The end-to-end loop is measured under the protocol of the device-resident one: from an idle device (the board's power governor
    # averages over ~1 s: straight after 25 back-to-back steps the SM clock is already capped, r02aw / r02bb), W warm-up steps, K timed.
    barrier()
    time.sleep(args.e2e_idle_s)
    for _ in range(W_):
        step_e2e()

    def timed_e2e(steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_e2e()
        e2e_losses.append(feed.drain())                          # the last step's loss is on the host before the clock stops
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    ms_e2e = timed_e2e(K)
    if not all(l == l and abs(l) < 1e30 for l in e2e_losses):
        raise SystemExit(f"[bench] non-finite loss in the end-to-end loop: {e2e_losses[-3:]}")
    value = B * world * K / (ms_total / 1e3)
    e2e_value = B * world * K / (ms_e2e / 1e3)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample (~10 s of CPU work on the box's cores): 3 timed steps of 8 patches (BASELINE.json configs[0] runs the
        # CPU case at batch 8) after 1 warm-up step; cfg4's model is 2.3x the work per patch, so it gets 4 patches
        cb = ref_batch(args, cfg)
        v, ms, cores, times = cpu_reference(cfg, 5, 2, cb, args.dist)
        cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{cb} patches/step x 5 timed steps (after 2 warm-up steps) of the {args.config} training step; value = patches / "
                              f"median step; oracle/torch_port.py eager fp32 (bit-equal to the imported reference on CPU) — the same "
                              f"function, batch and statistic as `bench.py --impl reference`",
                    "step_ms_min_median_max": [min(times), ms, max(times)]}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_, "extra_untimed_warmup_steps": extra_warm, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, cfg), "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": "per-step working set (>= 5 GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "conv_math": "cuDNN fp32 with TF32 allowed (PyTorch default, as the reference)",
                       "activation_layout": "NCHW" if args.nchw else "channels_last", "launch": graph_note,
                       "rgb_channel_padding": args.pad_rgb, "gdn": args.gdn, "fused_first_layer": bool(_layers.FUSE_FIRST_LAYER), "gemm_last_layer": bool(_layers.FAST_LAST_LAYER), "hyper_branch_on_side_stream": bool(_model.OVERLAP_HYPER_BRANCH),
                       "step_tail": "MS-SSIM scales + combination + clamp, loss tail, bias + ReLU, gradient pack, clip + Adam as sic kernels (DESIGN 5e)" if trainer.fused else "torch ops",
                       "gradient_buckets": [hi - lo for lo, hi, _, _ in trainer.buckets] if world > 1 else None,
                       "cudnn_benchmark": not args.no_cudnn_benchmark, "cudnn_benchmark_limit": torch.backends.cudnn.benchmark_limit},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4 * world, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": ms_e2e / K, "last_loss": e2e_losses[-1], "protocol": f"{args.e2e_idle_s} s idle, {W_} warm-up steps, {K} timed",
                    "how": "HostFedLoop: pinned batch -> copy stream -> landing buffer -> static input; loss -> pinned host, read one step late"},
            "gpu_launches": launches, "per_rank_ms_per_step": per_rank_ms or None, "clocks": clocks, "roofline": roof, "kernels": kernels, "cpu_baseline": cpu_base,
            "gpu_eager_baseline": gpu_eager,
            "allreduce_bytes_per_step": trainer.nbytes_allreduce if world > 1 else 0,
        }
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        # Round 1 left through os._exit: destroy_process_group() hung at 8 ranks while the captured graph (which holds the NCCL
        # kernels) was still alive.  Order matters: drop the graph first, then meet, then tear the communicator down.
        trainer.release_graph()
        del feed, run_static
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def kernel_rooflines(cfg, dev, channels_last=True, fused_first_layer=True):
    """CUDA-event timing of our own kernels at the sizes they have inside the step (largest GDN site, the latent), inputs
    larger than L2 or L2 flushed in between.  achieved = algorithmic bytes / time (GDN fwd 8 B/elem,
    bwd 12 B/elem; K1 fwd 12 B/elem broadcast, bwd 8 B/elem + 4 for the dense upstream of y_tilde)."""
    import torch
    from domain_specific_image_compression_b200 import functional as F
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def time_it(fn, reps=10, big=False):
        """big=True: the kernel streams >= 256 MB, twice the 126 MB L2 and walked front to back, so consecutive launches cannot
        reuse each other's lines and no flush is needed (B200_PROFILING.md: "either flush L2 or use inputs larger than L2").
        Flushing by WRITING 256 MB, as the small cases do, leaves the L2 full of dirty lines whose write-back (126 MB, ~8 % of a
        1.6 GB kernel) is then billed to the kernel under test: ncu, which invalidates instead, timed the same launches 7 % faster.
        The `reps` launches are queued back to back behind a ~1 ms device-side spin, so the host (45 us of Python per autograd
        call) is a full queue ahead and the launch-to-launch gap is the device's own, as inside the step's CUDA graph; one event
        pair around the batch, time / reps, median of 3 batches."""
        for _ in range(3):
            fn()
        if big:
            ts = []
            for _ in range(3):
                torch.cuda._sleep(2_000_000)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / reps)
            ts.sort()
            return ts[1] * 1e-3
        ts = []
        for _ in range(reps):
            flush.zero_()
            # a queued spin (~0.2 ms) lets the host run ahead: without it the GPU idles between e0 and the kernel while Python
            # (autograd dispatch, ~45 us for a backward) is still launching, and that idle time was being billed to the kernel
            torch.cuda._sleep(400_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2] * 1e-3

    B, N, M = cfg["batch"], cfg["N"], cfg["M"]
    out = {}
    x = torch.randn(B, N, 256, 256, device=dev)
    g = torch.randn_like(x)
    beta = torch.sqrt(torch.rand(N, device=dev) + 0.5).requires_grad_(True)
    w = torch.sqrt(torch.rand(N, 1, 1, 1, device=dev) * 0.3 + 0.01).requires_grad_(True)
    n = x.numel()
    t = time_it(lambda: F.gdn(x, beta, w, False), big=True)
    out["gdn_fwd"] = {"shape": list(x.shape), "bytes": 8 * n, "ms": t * 1e3, "gbs": 8 * n / t / 1e9}
    xr = x.clone().requires_grad_(True)
    y = F.gdn(xr, beta, w, False)
    t = time_it(lambda: torch.autograd.grad(y, (xr, beta, w), g, retain_graph=True), big=True)
    out["gdn_bwd"] = {"shape": list(x.shape), "bytes": 12 * n, "ms": t * 1e3, "gbs": 12 * n / t / 1e9}
    xc = x.contiguous(memory_format=torch.channels_last)
    gc = g.contiguous(memory_format=torch.channels_last)
    t = time_it(lambda: F.gdn(xc, beta, w, False), big=True)
    out["gdn_fwd_channels_last"] = {"shape": list(x.shape), "bytes": 8 * n, "ms": t * 1e3, "gbs": 8 * n / t / 1e9}
    xcr = xc.clone().requires_grad_(True)
    yc = F.gdn(xcr, beta, w, False)
    t = time_it(lambda: torch.autograd.grad(yc, (xcr, beta, w), gc, retain_graph=True), big=True)
    out["gdn_bwd_channels_last"] = {"shape": list(x.shape), "bytes": 12 * n, "ms": t * 1e3, "gbs": 12 * n / t / 1e9}
    del x, g, xr, y, xc, gc, xcr, yc
    # the 128^2 sites (two GDN + two IGDN per step): the largest GDN/IGDN sites of the step once the first layer is fused
    x = torch.randn(B, N, 128, 128, device=dev).contiguous(memory_format=torch.channels_last)
    g = torch.randn_like(x)
    n = x.numel()
    t = time_it(lambda: F.gdn(x, beta, w, False), big=8 * n >= (256 << 20))
    out["gdn_fwd_channels_last_128"] = {"shape": list(x.shape), "bytes": 8 * n, "ms": t * 1e3, "gbs": 8 * n / t / 1e9}
    xr = x.clone().requires_grad_(True)
    y = F.gdn(xr, beta, w, False)
    t = time_it(lambda: torch.autograd.grad(y, (xr, beta, w), g, retain_graph=True), big=12 * n >= (256 << 20))
    out["gdn_bwd_channels_last_128"] = {"shape": list(x.shape), "bytes": 12 * n, "ms": t * 1e3, "gbs": 12 * n / t / 1e9,
                                        "note": "both launches of the site's backward: gdn_bwd_nhwc_kernel + gdn_bwd_finalize_kernel"}
    # the streaming kernel on its own (sic_gdn_bwd_partials: the C-ABI half that launches only gdn_bwd_nhwc_kernel), at both sites
    import ctypes
    from domain_specific_image_compression_b200 import _lib as _L
    lib = _L.load()
    vp = lambda tt: ctypes.c_void_p(tt.data_ptr())
    for tag, xx, gg in (("gdn_bwd_nhwc_kernel_alone_128", x, g),):
        Bq, Cq, Hq, Wq = xx.shape
        ws = torch.zeros(lib.sic_gdn_bwd_workspace_bytes(Bq, Cq, Hq * Wq), dtype=torch.uint8, device=dev)
        dxq = torch.empty_like(xx)
        bq, wq = beta.detach(), w.detach().reshape(-1).contiguous()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        def main_only():
            rc = lib.sic_gdn_bwd_partials(vp(xx), None, vp(gg), vp(bq), vp(wq), Bq, Cq, Hq * Wq, 0, 1, vp(dxq), vp(ws), ws.numel(), st)
            assert rc == 0, lib.sic_last_error()
        t = time_it(main_only, big=12 * xx.numel() >= (256 << 20))
        out[tag] = {"shape": list(xx.shape), "bytes": 12 * xx.numel(), "ms": t * 1e3, "gbs": 12 * xx.numel() / t / 1e9,
                    "note": "gdn_bwd_nhwc_kernel only (dx + per-CTA partials), launched through sic_gdn_bwd_partials"}
        del ws, dxq
    del x, g, xr, y
    # N2: the fused first analysis layer (image in, y out / grad_y in): algorithmic bytes = the activation tensor once + the image
    try:
        img = torch.rand(B, 3, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
        w0 = (torch.randn(N, 3, 3, 3, device=dev) * 0.3).contiguous(memory_format=torch.channels_last)
        b0 = torch.randn(N, device=dev) * 0.2
        n = B * N * 256 * 256
        nb = 4 * (n + img.numel())
        t = time_it(lambda: F.conv0_gdn(img, w0, b0, beta, w), big=True)
        out["conv0_gdn_fwd_tcgen05"] = {"shape": [B, N, 256, 256], "bytes": nb, "ms": t * 1e3, "gbs": nb / t / 1e9}
        ps = [v.detach().clone().requires_grad_(True) for v in (w0, b0, beta, w)]
        y = F.conv0_gdn(img, *ps)
        g = torch.randn_like(y)
        t = time_it(lambda: torch.autograd.grad(y, ps, g, retain_graph=True), big=True)
        out["conv0_gdn_bwd_tcgen05"] = {"shape": [B, N, 256, 256], "bytes": nb, "ms": t * 1e3, "gbs": nb / t / 1e9,
                                        "note": "both launches of the backward (main kernel + fold of the per-CTA partials)"}
        del img, y, g
    except Exception as e:
        out["conv0_gdn_fwd_tcgen05"] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    # (e) step tail: clip + Adam on flat buffers (4 B read for the norm + 16 B read / 12 B written for the update, per parameter), at
    # the size of this model's flat buffer (L2 flushed between launches: 104 MB of state would otherwise stay resident) and at 64 M
    # parameters (1 GB of state: larger than L2); and the whole distortion term (3 scales forward + combination, then backward)
    try:
        n_model = 6_270_275 if N == 128 else 14_468_163                  # live parameters of the cfg2 / cfg4 model (FlatTrainer.flat)
        for tag, n_par, big in (("clip_adam_step_model_size", n_model, False), ("clip_adam_step_64M", 64 << 20, True)):
            p_, g_ = torch.randn(n_par, device=dev) * 1e-2, torch.randn(n_par, device=dev) * 1e-3
            m_, v_ = torch.zeros_like(p_), torch.zeros_like(p_)
            st_, nrm_, ws_ = torch.zeros((), device=dev), torch.zeros((), device=dev), F.clip_adam_workspace(n_par, dev)
            t = time_it(lambda: F.clip_adam_step(p_, g_, m_, v_, st_, nrm_, ws_, inv_world=1.0, clip=1.0, lr=1e-4, betas=(0.9, 0.999),
                                                 eps=1e-8, weight_decay=0.0), big=big)
            out[tag] = {"shape": [n_par], "bytes": 32 * n_par, "ms": t * 1e3, "gbs": 32 * n_par / t / 1e9,
                        "note": "both launches (grad_sumsq_kernel + adam_clip_kernel)"}
            del p_, g_, m_, v_
        from domain_specific_image_compression_b200 import losses as _losses
        xm = torch.rand(B, 3, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
        xr = (xm + 0.1 * torch.randn_like(xm)).requires_grad_(True)
        wm = torch.tensor([0.3, 0.5, 0.2], device=dev)
        t = time_it(lambda: _losses.multi_scale_ssim(xr, xm, 1.0, wm, clamp01=True))
        val = _losses.multi_scale_ssim(xr, xm, 1.0, wm, clamp01=True)
        tb = time_it(lambda: torch.autograd.grad(val, xr, retain_graph=True))
        out["msssim_distortion_fwd_bwd"] = {"shape": [B, 3, 256, 256], "ms_fwd": t * 1e3, "ms_bwd": tb * 1e3,
                                            "note": "clamp + 3 scales + combination: 4 launches forward, 3 backward; latency-sized (12.6 MB images)"}
        del xm, xr, val
    except Exception as e:
        out["clip_adam_step_model_size"] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    # likelihood kernel: the step's latent is tiny (launch-latency bound); the roofline figure is quoted on the top of the
    # BASELINE cfg5 sweep (128x128x320 latent, batch 16 = 84M elements, 1 GB of traffic)
    for tag, shape in (("k1_fwd_step", (B, M, 16, 16)), ("k1_fwd_sweep_top", (16, 320, 128, 128))):
        yl = torch.randn(*shape, device=dev) * 3
        sg = torch.exp(torch.randn(shape[0], shape[1], 1, 1, device=dev))
        nu = torch.exp(torch.randn(shape[0], shape[1], 1, 1, device=dev) + 1.5)
        t = time_it(lambda: F.bottleneck(yl, sg, nu, quant="noise"), big=yl.numel() * 12 >= (512 << 20))
        ne = yl.numel()
        out[tag] = {"shape": list(shape), "bytes": 12 * ne, "ms": t * 1e3, "gbs": 12 * ne / t / 1e9}
        if tag == "k1_fwd_sweep_top":
            # backward (8 B/elem + 4 for the dense upstream of y_tilde), spatial sigma/nu forward (20 B/elem) and the cdf_diff
            # likelihood (north_star's named kernel: SFU/issue-bound, so it is quoted in elements/s and MUFU thread-ops/s)
            yr, sr, nr = yl.clone().requires_grad_(True), sg.clone().requires_grad_(True), nu.clone().requires_grad_(True)
            yt, _, bits = F.bottleneck(yr, sr, nr, quant="noise")
            gb, gy = torch.ones_like(bits), torch.randn_like(yt)
            t = time_it(lambda: torch.autograd.grad((bits, yt), (yr, sr, nr), (gb, gy), retain_graph=True), big=True)
            out["k1_bwd_sweep_top"] = {"shape": list(shape), "bytes": 12 * ne, "ms": t * 1e3, "gbs": 12 * ne / t / 1e9}
            del yr, yt, gy
            ss, ns = sg.expand_as(yl).contiguous(), nu.expand_as(yl).contiguous()
            t = time_it(lambda: F.bottleneck(yl, ss, ns, quant="noise"), big=True)
            out["k1_fwd_spatial_sweep_top"] = {"shape": list(shape), "bytes": 20 * ne, "ms": t * 1e3, "gbs": 20 * ne / t / 1e9}
            del ss, ns
            t = time_it(lambda: F.bottleneck(yl, sg, nu, quant="noise", lik="cdf_diff"), reps=5, big=True)
            row = {"shape": list(shape), "bytes": 12 * ne, "ms": t * 1e3, "gbs": 12 * ne / t / 1e9, "bound": "sfu/issue (not hbm)",
                   "gelem_per_s": ne / t / 1e9}
            pipes = _committed("ncu_pipes_cdfdiff.json")
            if pipes:                                          # MUFU thread-ops per launch counted from SASS by ncu at this very shape
                e = pipes[0]
                if e.get("registers") == _regs("bottleneck_fwd_kernel<3,1,3,0>"):
                    peak_mufu = 148 * 16 * 1.965e9
                    row.update({"mufu_thread_ops_per_launch": e["mufu_thread_ops"], "mufu_ops_per_s": e["mufu_thread_ops"] / t,
                                "mufu_peak_ops_per_s": peak_mufu, "mufu_frac_of_peak": e["mufu_thread_ops"] / t / peak_mufu,
                                "warp_inst_per_launch": e["warp_inst"], "issue_frac_of_peak": e["warp_inst"] / t / (148 * 4 * 1.965e9),
                                "counts_source": "profiles/ncu_pipes_cdfdiff.json (ncu SASS page, same register allocation)"})
            out["k1_fwd_cdfdiff_sweep_top"] = row
        del yl
    # G3 (north_star's tensor-core contraction; not on the default step, which runs the reference's diagonal GDN): pipelined
    # tcgen05 kernel at the largest site shape, C = 128 (the N=128 model) and C = 192 (the N=192 model of cfg4)
    for Cd, Bd in ((128, 16), (192, 8)):
        xd = torch.randn(Bd, Cd, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
        gm = torch.sqrt(torch.rand(Cd, Cd, device=dev) * 0.02 + torch.eye(Cd, device=dev) * 0.1 + 2.0 ** -18)
        bd = torch.sqrt(torch.rand(Cd, device=dev) + 0.5)
        t = time_it(lambda: F.gdn_dense(xd, bd, gm, False), big=True)
        nd = xd.numel()
        out[f"gdn_dense_fwd_tcgen05_c{Cd}"] = {"shape": list(xd.shape), "bytes": 8 * nd, "ms": t * 1e3, "gbs": 8 * nd / t / 1e9,
                                               "tf32_mma_tflops": 4.0 * Cd * nd / t / 1e12}
        xr, br, gr = xd.clone().requires_grad_(True), bd.clone().requires_grad_(True), gm.clone().requires_grad_(True)
        yv = F.gdn_dense(xr, br, gr, False)
        go = torch.randn_like(yv)
        t = time_it(lambda: torch.autograd.grad(yv, (xr, br, gr), go, retain_graph=True), reps=5, big=True)
        out[f"gdn_dense_bwd_tcgen05_c{Cd}"] = {"shape": list(xd.shape), "bytes": 12 * nd, "ms": t * 1e3, "gbs": 12 * nd / t / 1e9,
                                               "note": "dx, d(beta), d(gamma): all launches of the backward"}
        del xd, xr, yv, go
    for v in out.values():
        if "gbs" in v:
            v["frac_of_hbm_peak"] = v["gbs"] / peak
    # The dominant own kernel of the step is the GDN backward at the largest site, in the layout the step actually runs
    # (channels_last by default -> gdn_bwd_nhwc_kernel; --nchw -> gdn_bwd_kernel).  `traffic` = dram__bytes_read.sum +
    # dram__bytes_write.sum of that kernel at this shape, looked up in profiles/ncu_traffic.json (written by scripts/ncu_traffic.py
    # from a committed ncu --set full capture) and used ONLY if the capture's registers/thread equal those of the loaded library.
    if channels_last and fused_first_layer:
        dom, kid = out["gdn_bwd_nhwc_kernel_alone_128"], "gdn_bwd_nhwc_kernel<0>"
        kname = ("gdn_bwd_nhwc_kernel (GDN/IGDN backward, channels_last; 12 launches and the largest share of the step among our kernels: "
                 "profiles/r02ax_ncu_launches_bench_step.txt) at its largest site in the step, 128^2 (the 256^2 site is inside the fused "
                 "first-layer kernel; this kernel at 256^2: kernels.gdn_bwd_channels_last)")
    elif channels_last:
        dom, kid = out["gdn_bwd_channels_last"], "gdn_bwd_nhwc_kernel<0>"
        kname = "gdn_bwd_nhwc_kernel (GDN backward, channels_last) at the largest site of the step"
    else:
        dom, kid = out["gdn_bwd"], "gdn_bwd_kernel<0,1>"
        kname = "gdn_bwd_kernel (GDN backward, NCHW) at the largest site of the step"
    traffic, tsrc = None, None
    regs = _regs(kid)
    for e in _committed("ncu_traffic.json") or []:
        if e["kernel"] == kid and e["shape"] == dom["shape"] and e["registers"] == regs:
            traffic = e["dram_read_bytes"] + e["dram_write_bytes"]
            tsrc = f"profiles/ncu_traffic.json <- {e['capture']} ({kid}, {regs} registers/thread: matches the loaded library)"
    roof = {"kernel": kname, "bound": "hbm", "achieved": dom["gbs"], "peak": peak,
            "unit": "GB/s", "frac": dom["gbs"] / peak, "traffic": traffic, "traffic_source": tsrc, "registers_per_thread": regs,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": dom["bytes"], "ms_per_launch": dom["ms"],
            "site_backward_ms_incl_fold_launch": out.get("gdn_bwd_channels_last_128", {}).get("ms") if (channels_last and fused_first_layer) else None,
            "timing": "CUDA events on the launching stream around 10 back-to-back launches of this kernel (through the C ABI: "
                      "sic_gdn_bwd_partials when the first layer is fused, else the site's whole backward incl. the per-channel fold "
                      "launch), queued behind a 1 ms device-side spin so the host is a full queue ahead; working set >= 3 x the 126 MB "
                      "L2, walked front to back: no flush needed, none done; time / 10, median of 3 batches"}
    return roof, out


def _committed(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except (OSError, ValueError):
        return None


def _regs(kernel_id):
    from domain_specific_image_compression_b200 import _lib
    r = _lib.load().sic_kernel_registers(kernel_id.encode())
    return r if r > 0 else None


def main():
    args = parse()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on init), so
    # route fd 1 to stderr for the whole run and keep a private handle on the real stdout for the result line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
