"""Data-parallel training step: one process per GPU, flat gradient buffer, bucketed NCCL all-reduce overlapped with backward.

The reference trains on a single GPU (train.py:191-204: zero_grad, forward 'noise', loss, backward, clip 1.0, Adam) and has
no distributed code at all.  Patches are independent, so the B200 design is pure data parallelism over the
NVSwitch domain: every rank holds a full replica (26-60 MB of weights), runs the reference's step on its own patches, and
the only exchange is the gradient all-reduce.  The parameters live in ONE flat fp32 buffer (the modules' tensors are views
into it) so that
  - the gradients are reduced in place in ONE flat buffer (25.9 MB for N=128,M=192; 59.8 MB for N=192,M=320) cut into a few
    contiguous buckets in reverse parameter order; each bucket's ncclAllReduce is issued (asynchronously, on NCCL's stream)
    from an autograd hook the moment its last gradient exists, so it runs under the rest of the backward pass — g_s and the
    hyper networks finish first and the 256^2 / 128^2 analysis sites that follow take milliseconds.  Only the last, small
    bucket (the first analysis layers) is exposed.  Round 1 issued one all-reduce after backward() had returned: 7.04 -> 7.20
    ms per step at 8 GPUs with nothing left to hide it under;
  - global-norm clipping and Adam are two launches on the flat buffer (csrc/train_step.cu: one pass for the norm, one for the
    update with the clip coefficient folded in) instead of ~80 per-tensor launch groups;
  - the dead parameters (per GDN site: the CxC `gamma` on the reference's diagonal path, layers.py:13; `gamma_conv.weight`
    when the site runs dense) are simply left out of the bucket — plain
    DistributedDataParallel would need find_unused_parameters for them.
Clipping is applied AFTER the all-reduce so that all ranks scale identically.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional

import torch
import torch.distributed as dist


# SIC_DIAG_NO_ALLREDUCE=1: timing diagnosis only (how much of a multi-GPU step is the collective?) - replicas diverge, never train so
import os as _os
_DIAG_NO_ALLREDUCE = bool(_os.environ.get("SIC_DIAG_NO_ALLREDUCE"))


def _dead_parameter_names(module: torch.nn.Module) -> set:
    """Parameters that never receive a gradient, decided per GDN site: the diagonal path (the reference's, layers.py:21) uses
    `gamma_conv.weight` and leaves the stored CxC `gamma` (layers.py:13) untouched; GDN(dense=True) is the other way round."""
    from .layers import GDN
    dead = set()
    for prefix, m in module.named_modules():
        if isinstance(m, GDN):
            dot = prefix + "." if prefix else ""
            dead.add(dot + ("gamma_conv.weight" if m.dense else "gamma"))
    return dead


def _flat_in_param_order(g: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """Flatten a gradient in the MEMORY order of its parameter (the order the flat buffer stores the parameter in)."""
    if _is_dense_permutation(p):
        return g.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1).reshape(-1)
    return g.reshape(-1)


def _is_dense_permutation(p: torch.Tensor) -> bool:
    """True for non-contiguous tensors that still cover numel() distinct elements of one block (e.g. channels_last)."""
    return (not p.is_contiguous()) and p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last)


def _is_capturing(stream: "torch.cuda.Stream") -> bool:
    with torch.cuda.stream(stream):
        return torch.cuda.is_current_stream_capturing()


class FlatTrainer:
    def __init__(self, module: torch.nn.Module, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, grad_clip: float = 1.0, process_group=None, fused: Optional[bool] = None,
                 exclude: Iterable[str] = (), bucket_bytes: Optional[int] = None):
        self.module = module
        self.grad_clip = grad_clip
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        if bucket_bytes is None:               # one process: nothing to overlap, one pack launch at the end of backward() instead of five
            bucket_bytes = (6 << 20) if self.world > 1 else (1 << 62)
        dead = _dead_parameter_names(module) | set(exclude)      # `exclude`: further parameter names the step never touches
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad and n not in dead]
        self.names = [n for n, _ in named]
        self.live: List[torch.nn.Parameter] = [p for _, p in named]
        total = sum(p.numel() for p in self.live)
        ref = self.live[0]
        self.flat = torch.empty(total, dtype=ref.dtype, device=ref.device)
        offsets, off = [], 0
        for p in self.live:                    # re-seat every parameter as a view into the flat buffer
            n = p.numel()
            view = self.flat[off:off + n].as_strided(p.size(), p.stride()) if _is_dense_permutation(p) else self.flat[off:off + n].view_as(p)
            view.copy_(p.detach())             # keeps a channels_last weight channels_last inside the flat buffer
            p.data = view
            offsets.append(off)
            off += n
        self._offsets = offsets
        self.flat.requires_grad_(True)
        self.flat_grad = torch.zeros_like(self.flat)              # persistent: what the all-reduces and Adam work on
        if self.world > 1:                     # identical replicas: rank 0's weights win
            dist.broadcast(self.flat.detach(), src=0, group=process_group)
        if fused is None:
            fused = self.flat.is_cuda
        # fused (the default on a GPU): clip + Adam are the two launches of csrc/train_step.cu on the flat buffers; the update
        # counter and the norm live on the device, so the whole step can be recorded into a CUDA graph.  fused=False keeps
        # torch.optim.Adam on the flat buffer (the CPU / gloo tests of the host logic, and the comparison in the GPU test).
        self.fused = bool(fused)
        self.hyper = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.opt = None
        if self.fused:
            from . import functional as F_sic
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
            self.step_count = torch.zeros((), dtype=torch.float32, device=self.flat.device)
            self.grad_norm = torch.zeros((), dtype=torch.float32, device=self.flat.device)
            self._opt_ws = F_sic.clip_adam_workspace(total, self.flat.device)
        else:
            self.opt = torch.optim.Adam([self.flat], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, fused=False,
                                        capturable=bool(self.flat.is_cuda))
        self.graph = None
        self.nbytes_allreduce = total * self.flat.element_size()
        # Buckets: contiguous slices of the flat buffer, formed walking the parameters BACKWARDS (the order backward() roughly
        # produces gradients in) and closed once they hold bucket_bytes.  (lo, hi, first_param, last_param + 1)
        self.buckets = []
        hi_i, acc = len(self.live), 0
        for i in range(len(self.live) - 1, -1, -1):
            acc += self.live[i].numel() * self.flat.element_size()
            if acc >= bucket_bytes or i == 0:
                lo = offsets[i]
                hi = offsets[hi_i - 1] + self.live[hi_i - 1].numel()
                self.buckets.append((lo, hi, i, hi_i))
                hi_i, acc = i, 0
        self._bucket_of = {}
        for k, (_, _, a, b) in enumerate(self.buckets):
            for i in range(a, b):
                self._bucket_of[id(self.live[i])] = k
        self._pending = [0] * len(self.buckets)
        self._works = []
        self._armed = False
        self._main_stream = None
        self.fire_order: List[int] = []        # bucket indices in the order their all-reduce was issued in the last step
        for p in self.live:
            p.register_post_accumulate_grad_hook(self._on_grad)

    # ------------------------------------------------------------------------------------------------- gradient buckets
    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not self._armed:
            return
        k = self._bucket_of[id(p)]
        self._pending[k] -= 1
        if self._pending[k] == 0:
            self._fire(k)

    def _fire(self, k: int) -> None:
        """Every gradient of bucket k exists: pack them into the bucket's slice of the flat buffer and start its all-reduce."""
        lo, hi, a, b = self.buckets[k]
        if self.flat_grad.is_cuda:
            # The gradients of one bucket may have been produced on the step's own stream or on the model's side stream
            # (model.OVERLAP_HYPER_BRANCH), and this hook runs on whichever of the two produced the last one: wait for both.
            from .model import _side_streams
            # While a CUDA graph is being recorded only streams that were forked into the capture may be waited for: a side stream
            # that exists (another model used it earlier) but did no work in this step is not part of it.
            cur = torch.cuda.current_stream(self.flat_grad.device)
            capturing = torch.cuda.is_current_stream_capturing()
            for st in (self._main_stream, _side_streams.get(self.flat_grad.device.index)):
                if st is not None and st != cur and (not capturing or _is_capturing(st)):
                    cur.wait_stream(st)
        grads = [_flat_in_param_order(self.live[i].grad, self.live[i]) for i in range(a, b)]
        if self.flat_grad.is_cuda and self.fused:
            from . import functional as F_sic
            keep = [g for g in grads if g.numel() > 0]
            F_sic.pack_flat([g.contiguous() for g in keep], [self._offsets[i] for i, g in zip(range(a, b), grads) if g.numel() > 0],
                            self.flat_grad)
        else:
            torch.cat(grads, out=self.flat_grad[lo:hi])
        self.fire_order.append(k)
        if self.world > 1 and not _DIAG_NO_ALLREDUCE:   # asynchronous: runs on the collective's own stream under the rest of backward()
            self._works.append(dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def pack_grads(self) -> torch.Tensor:
        """The flat gradient once backward() has run (buckets already packed — and being reduced — by the hooks)."""
        missing = [self.names[i] for k, (_, _, a, b) in enumerate(self.buckets) if self._pending[k] > 0
                   for i in range(a, b) if self.live[i].grad is None]
        if missing:
            raise RuntimeError(f"parameters without gradient after backward: {missing[:4]}...")
        return self.flat_grad

    def reduce_clip_step(self, flat_grad: torch.Tensor) -> torch.Tensor:
        """wait for the bucket all-reduces (mean) -> global-norm clip (torch.nn.utils.clip_grad_norm_ semantics, train.py:200-202)
        -> Adam.  Returns the pre-clip global gradient norm (device scalar, no host sync)."""
        for w in self._works:                  # stream-level wait (the host does not block on CUDA)
            w.wait()
        self._works = []
        if self.fused:
            from . import functional as F_sic
            F_sic.clip_adam_step(self.flat.detach(), flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.grad_norm,
                                 self._opt_ws, inv_world=1.0 / self.world, clip=self.grad_clip or 0.0, **self.hyper)
            return self.grad_norm
        if self.world > 1:
            flat_grad.mul_(1.0 / self.world)
        norm = torch.linalg.vector_norm(flat_grad)
        if self.grad_clip and self.grad_clip > 0:
            flat_grad.mul_(torch.clamp(self.grad_clip / (norm + 1e-6), max=1.0))
        self.flat.grad = flat_grad
        self.opt.step()
        return norm

    def step(self, loss_closure: Callable[[], torch.Tensor]) -> torch.Tensor:
        """loss_closure runs forward + loss on this rank's patches and returns the scalar loss."""
        for p in self.live:
            p.grad = None                      # zero_grad(set_to_none=True), train.py:195
        for k, (_, _, a, b) in enumerate(self.buckets):
            self._pending[k] = b - a
        self._works, self.fire_order, self._armed = [], [], True
        self._main_stream = torch.cuda.current_stream(self.flat_grad.device) if self.flat_grad.is_cuda else None
        try:
            loss = loss_closure()
            loss.backward()
        finally:
            self._armed = False
        self.reduce_clip_step(self.pack_grads())
        return loss.detach()

    def capture(self, loss_closure: Callable[[], torch.Tensor], warmup: int = 3) -> Callable[[], torch.Tensor]:
        """Record zero_grad + forward + loss + backward + the bucket all-reduces + clip + Adam into ONE CUDA graph and return a
        replay function.  The closure must read its batch from a static tensor (copy the next batch into it before each
        replay).  The in-kernel Philox offset advances on the device, so every replay draws fresh quantisation noise.
        Launch-bound inner loops are what graphs are for: the step is ~600 small launches next to a dozen large ones."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):            # cuDNN autotune, workspaces, Philox state: all allocated before capture
                self.step(loss_closure)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        for p in self.live:
            p.grad = None
        with torch.cuda.graph(graph):
            static_loss = self.step(loss_closure)
        self.graph = graph

        def replay() -> torch.Tensor:
            graph.replay()
            return static_loss
        return replay

    def release_graph(self) -> None:
        """Drop the captured step (its NCCL kernels reference the communicator): call before destroy_process_group()."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None
            for p in self.live:
                p.grad = None


class HostFedLoop:
    """Feeds a (captured or eager) step from pinned HOST batches without stalling the device on the copies.

    Per step: the batch goes host -> device on its own copy stream into one of two landing buffers while the previous step is still
    computing; the step stream waits for that copy, moves the 12.6 MB batch into the step's static input (device-to-device),
    runs the step; the scalar loss goes device -> host on a second copy stream and is handed back ONE STEP LATE, so the host never
    blocks on the step it has just enqueued.  Every step still pays its own H2D and D2H; they are simply no longer serialised
    with the kernels (round 1: blocking copy_ before and after the replay, 0.14 ms/step on one GPU, 0.50 ms at eight)."""

    def __init__(self, run_step: Callable[[], torch.Tensor], x_static: torch.Tensor):
        self.run_step, self.x_static = run_step, x_static
        # two copy streams: on ONE in-order stream the next batch's H2D would queue behind the previous loss's D2H, which waits
        # for the previous step to finish - the H2D then ran between the steps instead of under them (measured: +0.21 ms/step)
        self.h2d_stream = torch.cuda.Stream(device=x_static.device)
        self.d2h_stream = torch.cuda.Stream(device=x_static.device)
        self.land = [torch.empty_like(x_static) for _ in range(2)]
        self.loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.h2d_done = [torch.cuda.Event() for _ in range(2)]
        self.d2h_done = [torch.cuda.Event() for _ in range(2)]
        self.step_done = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self._staged = False

    def stage(self, x_host: torch.Tensor) -> None:
        """Start the host->device copy of the batch of the NEXT step() call."""
        k = self.i % 2
        with torch.cuda.stream(self.h2d_stream):
            self.h2d_stream.wait_event(self.consumed[k])           # the step that last read this landing buffer has moved it on
            self.land[k].copy_(x_host, non_blocking=True)
            self.h2d_done[k].record(self.h2d_stream)
        self._staged = True

    def step(self, x_host_next: Optional[torch.Tensor] = None) -> Optional[float]:
        """Runs one step on the staged batch, stages `x_host_next` for the following call, returns the PREVIOUS step's loss."""
        if not self._staged:
            raise RuntimeError("HostFedLoop.step: call stage(batch) first")
        k = self.i % 2
        main = torch.cuda.current_stream(self.x_static.device)
        main.wait_event(self.h2d_done[k])
        if self.i >= 1:
            main.wait_event(self.d2h_done[(self.i - 1) % 2])        # the previous loss has left the static tensor the step rewrites
        self.x_static.copy_(self.land[k], non_blocking=True)
        self.consumed[k].record(main)
        loss = self.run_step()
        self.step_done[k].record(main)
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(self.step_done[k])
            self.loss_host[k].copy_(loss, non_blocking=True)
            self.d2h_done[k].record(self.d2h_stream)
        self.i += 1
        self._staged = False
        if x_host_next is not None:
            self.stage(x_host_next)
        if self.i >= 2:
            prev = (self.i - 2) % 2
            self.d2h_done[prev].synchronize()                       # the loss of the step before the one just enqueued
            return float(self.loss_host[prev])
        return None

    def drain(self) -> float:
        """Loss of the last step (blocks until it is on the host)."""
        last = (self.i - 1) % 2
        self.d2h_done[last].synchronize()
        return float(self.loss_host[last])
