"""Data-parallel training step: one process per GPU, flat gradient bucket, one NCCL all-reduce per step.

The reference trains on a single GPU (train.py:191-204: zero_grad, forward 'noise', loss, backward, clip 1.0, Adam) and has
no distributed code at all.  Patches are independent, so the B200 design is pure data parallelism over the
NVSwitch domain: every rank holds a full replica (26-60 MB of weights), runs the reference's step on its own patches, and
the only exchange is the gradient all-reduce.  The parameters live in ONE flat fp32 buffer (the modules' tensors are views
into it) so that
  - the all-reduce is a single ncclAllReduce over 25.9 MB (N=128,M=192) / 59.8 MB (N=192,M=320): launch-latency bound on
    NVLink 5, in-switch reduction (NVLS) when NCCL enables it;
  - global-norm clipping and Adam are three launches on the flat buffer instead of ~80 per-tensor launch groups;
  - the dead parameters (per GDN site: the CxC `gamma` on the reference's diagonal path, layers.py:13; `gamma_conv.weight`
    when the site runs dense) are simply left out of the bucket — plain
    DistributedDataParallel would need find_unused_parameters for them.
Clipping is applied AFTER the all-reduce so that all ranks scale identically.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional

import torch
import torch.distributed as dist


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


class FlatTrainer:
    def __init__(self, module: torch.nn.Module, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, grad_clip: float = 1.0, process_group=None, fused: Optional[bool] = None,
                 exclude: Iterable[str] = ()):
        self.module = module
        self.grad_clip = grad_clip
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        dead = _dead_parameter_names(module) | set(exclude)      # `exclude`: further parameter names the step never touches
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad and n not in dead]
        self.names = [n for n, _ in named]
        self.live: List[torch.nn.Parameter] = [p for _, p in named]
        total = sum(p.numel() for p in self.live)
        ref = self.live[0]
        self.flat = torch.empty(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.live:                    # re-seat every parameter as a view into the flat buffer
            n = p.numel()
            view = self.flat[off:off + n].as_strided(p.size(), p.stride()) if _is_dense_permutation(p) else self.flat[off:off + n].view_as(p)
            view.copy_(p.detach())             # keeps a channels_last weight channels_last inside the flat buffer
            p.data = view
            off += n
        self.flat.requires_grad_(True)
        if self.world > 1:                     # identical replicas: rank 0's weights win
            dist.broadcast(self.flat.detach(), src=0, group=process_group)
        if fused is None:
            fused = self.flat.is_cuda
        # capturable: the step counter lives on the device, so the whole step can be recorded into a CUDA graph
        self.opt = torch.optim.Adam([self.flat], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, fused=fused,
                                    capturable=bool(self.flat.is_cuda))
        self.graph = None
        self.nbytes_allreduce = total * self.flat.element_size()

    def pack_grads(self) -> torch.Tensor:
        missing = [n for n, p in zip(self.names, self.live) if p.grad is None]
        if missing:
            raise RuntimeError(f"parameters without gradient after backward: {missing[:4]}...")
        return torch.cat([_flat_in_param_order(p.grad, p) for p in self.live])

    def reduce_clip_step(self, flat_grad: torch.Tensor) -> torch.Tensor:
        """all-reduce (mean) -> global-norm clip (torch.nn.utils.clip_grad_norm_ semantics, train.py:200-202) -> Adam.
        Returns the pre-clip global gradient norm (device scalar, no host sync)."""
        if self.world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
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
        loss = loss_closure()
        loss.backward()
        self.reduce_clip_step(self.pack_grads())
        return loss.detach()

    def capture(self, loss_closure: Callable[[], torch.Tensor], warmup: int = 3) -> Callable[[], torch.Tensor]:
        """Record zero_grad + forward + loss + backward + all-reduce + clip + Adam into ONE CUDA graph and return a
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
